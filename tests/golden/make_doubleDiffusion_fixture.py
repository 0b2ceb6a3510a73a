"""Builds tests/golden/doubleDiffusion_reference.npz from the REFERENCE's own test fixture samples/doubleDiffusion.

Run in the build container only (reads /root/reference); the .npz travels.

Sources: samples/doubleDiffusion/checkpoint_end.start (saturated double-diffusive convection: thermal AND compositional
buoyancy, l_max = 64 with minc = 4, n_r_max = 33 / n_cheb_max = 31, read with magic_b200.checkpoint), the first six rows of
reference.out (e_kin.TAG of the Chebyshev run of input.nml: 25 steps of the IMEX Runge-Kutta scheme BPR353 with dt = 3e-4,
logged every 5 steps; the following rows belong to the finite-difference run of input_FD.nml) and the values of input.nml.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from magic_b200.checkpoint import read_checkpoint  # noqa: E402

REF = "/root/reference/samples/doubleDiffusion"

ck = read_checkpoint(os.path.join(REF, "checkpoint_end.start"))
print(ck.version, ck.family, ck.l_max, ck.trunc, list(ck.fields), ck.params, ck.dt, ck.rscheme)
e_kin = np.loadtxt(os.path.join(REF, "reference.out"))[:6]
np.savez_compressed(os.path.join(HERE, "doubleDiffusion_reference.npz"), e_kin=e_kin, time=ck.time, radius=ck.r,
                    **{k: v for k, v in ck.fields.items()}, n_log_step=5, n_r_max=33, n_cheb_max=31, n_phi_tot=192, minc=4,
                    ra=4.8e4, raxi=1.2e5, ek=1e-3, pr=0.3, sc=3.0, radratio=0.35, dtmax=3e-4, alpha=0.6, ktopv=2, kbotv=2,
                    courfac=0.8, alffac=0.35)
print(e_kin)
