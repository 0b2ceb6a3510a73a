"""Extracts the golden vectors of samples/precession from the reference tree (run in the build container).

reference.out is the e_kin.TAG series MagIC's autotest compares against at rtol 1e-8 (samples/precession/unitTest.py): 200
CNAB2 steps of dt=1e-5 of a precessing shell (Po=-0.01, 23.5 degrees, Ek=1e-3, no buoyancy, no field, rigid walls) started
from rest, l_max=42 with m_max=5, n_r_max=n_cheb_max=49, logged every 10 steps (21 rows).  The values of input.nml the host
restatement needs are stored next to it.
"""
import os

import numpy as np

REF = "/root/reference/samples/precession"
HERE = os.path.dirname(os.path.abspath(__file__))

e_kin = np.loadtxt(os.path.join(REF, "reference.out"))
np.savez_compressed(os.path.join(HERE, "precession_reference.npz"), e_kin=e_kin, n_log_step=10, n_r_max=49, n_cheb_max=49,
                    l_max=42, m_max=5, minc=1, ra=0.0, ek=1e-3, pr=1.0, prmag=5.0, radratio=0.35, po=-1.0e-2, prec_angle=23.5,
                    dtmax=1e-5, alpha=0.6, ktopv=2, kbotv=2, courfac=2.5, alffac=1.0, intfac=3.0e-2)
print(e_kin.shape)
