"""Extracts the golden vectors of samples/varCond from the reference tree (run in the build container).

reference.out / referenceMag.out are the e_kin / e_mag_oc series MagIC's autotest compares against at rtol 1e-8
(samples/varCond/unitTest.py): an ANELASTIC dynamo (N_rho = 1, polytropic index 2, gravity ~ r) with a radially varying
electrical conductivity (nVarCond = 2), a conducting, non-rotating inner core (kbotb = 3, sigma_ratio = 1), stress-free
outer and rigid inner wall, l_correct_AMz / AMe; n_phi_tot = 96 -> l_max = 32, n_r_max = 49 with n_cheb_max = 47, radratio
0.2, CNAB2 with dt = 5e-6 from init_s1 = 505 / init_b1 = 3, 500 steps logged every 10 (51 rows).
"""
import os

import numpy as np

REF = "/root/reference/samples/varCond"
HERE = os.path.dirname(os.path.abspath(__file__))

e_kin = np.loadtxt(os.path.join(REF, "reference.out"))
e_mag = np.loadtxt(os.path.join(REF, "referenceMag.out"))
np.savez_compressed(os.path.join(HERE, "varCond_reference.npz"), e_kin=e_kin, e_mag_oc=e_mag, n_log_step=10, n_r_max=49,
                    n_cheb_max=47, n_r_ic_max=17, n_cheb_ic_max=15, n_phi_tot=96, minc=1, ra=1.5e7, ek=1e-4, pr=1.0, prmag=1.0,
                    strat=1.0, polind=2.0, radratio=0.2, g0=0.0, g1=1.0, g2=0.0, dtmax=5e-6, alpha=0.6, init_s1=505, amp_s1=0.1,
                    init_b1=3, amp_b1=5.0, courfac=2.5, alffac=1.0, sigma_ratio=1.0, ktopv=1, kbotv=2, kbotb=3, nVarCond=2,
                    con_DecRate=9.0, con_RadRatio=0.8, con_LambdaMatch=0.5, con_LambdaOut=0.1, con_FuncWidth=0.25)
print(e_kin.shape, e_mag.shape)
