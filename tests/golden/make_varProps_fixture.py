"""Extracts the golden vectors of samples/varProps (Chebyshev stage) from the reference tree (run in the build container).

reference.out concatenates the e_kin.TAG series of three runs of the same anelastic case (inputCheb.nml, inputMap.nml,
inputFD.nml; samples/varProps/unitTest.py, rtol 1e-8); the first 26 rows are the Chebyshev run restated here: N_rho = 3,
polytropic index 2, gravity ~ 1/r^2, stress-free walls, kinematic viscosity and thermal diffusivity proportional to
rho^-1/2 (nVarVisc = nVarDiff = 2, difExp = -0.5), n_phi_tot = 96 -> l_max = 32, n_r_max = n_cheb_max = 33, 250 CNAB2 steps of
1e-4 from init_s1 = 707, logged every 10 steps.
"""
import os

import numpy as np

REF = "/root/reference/samples/varProps"
HERE = os.path.dirname(os.path.abspath(__file__))

e_kin = np.loadtxt(os.path.join(REF, "reference.out"))[:26]
np.savez_compressed(os.path.join(HERE, "varProps_reference.npz"), e_kin=e_kin, n_log_step=10, n_r_max=33, n_cheb_max=33,
                    n_phi_tot=96, minc=1, ra=8.0e4, ek=1e-3, pr=1.0, prmag=5.0, strat=3.0, polind=2.0, radratio=0.35, g0=0.0,
                    g1=0.0, g2=1.0, dtmax=1e-4, alpha=0.6, init_s1=707, amp_s1=0.01, ktopv=1, kbotv=1, courfac=2.5, alffac=1.0,
                    nVarDiff=2, nVarVisc=2, difExp=-0.5)
print(e_kin.shape, e_kin[:3, :3])
