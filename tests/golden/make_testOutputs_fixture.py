"""Extracts the golden vectors of samples/testOutputs from the reference tree (run in the build container).

samples/testOutputs/reference.out is the concatenation e_kin, e_mag_oc, e_mag_ic, dipole, heat, par, power, u_square, helicity,
hemi (unitTest.py:67) of a 100-step run that MagIC's autotest compares at rtol 1e-8: a weakly stratified ANELASTIC dynamo
(strat = 0.1, polytropic index 2, gravity ~ r), rigid insulating walls, n_phi_tot = 256 -> l_max = 85, n_r_max = 73 with
n_cheb_max = 71, CNAB2 with dt = 1e-4 from init_s1 = 404 (amp 0.01) / init_b1 = 3, logged every 10 steps, with l_hel, l_hemi,
l_power and l_RMS on.  helicity.TAG, hemi.TAG and the viscous-dissipation column of power.TAG are radial integrals of what
get_helicity / get_hemi / get_visc_heat sum on the grid inside the radial loop (rIter.f90:320-342).
"""
import os

import numpy as np

REF = "/root/reference/samples/testOutputs"
HERE = os.path.dirname(os.path.abspath(__file__))

rows = [np.array(l.split(), dtype=float) for l in open(os.path.join(REF, "reference.out")) if l.strip()]
sec = {}
names = [("e_kin", 11), ("e_mag_oc", 11), ("e_mag_ic", 11), ("dipole", 11), ("heat", 11), ("par", 11), ("power", 10), ("u_square", 11),
         ("helicity", 11), ("hemi", 11)]
i = 0
for nm, n in names:
    sec[nm] = np.array(rows[i:i + n])
    i += n
assert i == len(rows) == 109
assert sec["helicity"].shape == (11, 9) and sec["hemi"].shape == (11, 8) and sec["power"].shape == (10, 11)
np.savez_compressed(os.path.join(HERE, "testOutputs_reference.npz"), n_log_step=10, n_r_max=73, n_cheb_max=71, n_phi_tot=256, minc=1,
                    ra=3.0e5, ek=1e-3, pr=1.0, prmag=5.0, strat=0.1, polind=2.0, radratio=0.35, g0=0.0, g1=1.0, g2=0.0, dtmax=1e-4,
                    alpha=0.6, init_s1=404, amp_s1=0.01, init_b1=3, amp_b1=5.0, courfac=2.5, alffac=1.0, ktopv=2, kbotv=2,
                    **{k: sec[k] for k in ("e_kin", "e_mag_oc", "power", "helicity", "hemi")})
print({k: v.shape for k, v in sec.items()})
