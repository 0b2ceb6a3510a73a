"""Extracts the golden vectors of samples/dynamo_benchmark_condICrotIC from the reference tree (run in the build container).

reference.out / referenceMag.out are the e_kin / e_mag_oc series MagIC's autotest compares against at rtol 1e-8
(samples/dynamo_benchmark_condICrotIC/unitTest.py): the Christensen benchmark with a finitely CONDUCTING (sigma_ratio = 1,
kbotb = 3) and freely ROTATING (nRotIC = 1) inner core, rigid walls, n_phi_tot = 48 -> l_max = 16, n_r_max = 33 with
n_cheb_max = 31, inner core n_r_ic_max = 17 with n_cheb_ic_max = 15, CNAB2 with dt = 1e-4 from the start fields of
init_s1 = 404 / init_b1 = 3, logged every step.  Rows 0..1000 are the first run (tag "start"); rows 1001..1101 are the
restarted run of input_restart.nml (stress-free walls, l_correct_AMz / AMe: nonlinear magnetic boundary condition at the ICB).  All 1102 rows are kept.
"""
import os

import numpy as np

REF = "/root/reference/samples/dynamo_benchmark_condICrotIC"
HERE = os.path.dirname(os.path.abspath(__file__))

e_kin = np.loadtxt(os.path.join(REF, "reference.out"))[:1102]
e_mag = np.loadtxt(os.path.join(REF, "referenceMag.out"))[:1102]
np.savez_compressed(os.path.join(HERE, "condICrotIC_reference.npz"), e_kin=e_kin, e_mag_oc=e_mag, n_r_max=33, n_cheb_max=31,
                    n_r_ic_max=17, n_cheb_ic_max=15, n_phi_tot=48, minc=1, ra=1.1e5, ek=1e-3, pr=1.0, prmag=5.0, radratio=0.35,
                    dtmax=1e-4, alpha=0.6, init_s1=404, amp_s1=0.1, init_b1=3, amp_b1=5.0, courfac=2.5, alffac=1.0,
                    sigma_ratio=1.0, nRotIC=1, ktopv=2, kbotv=2, kbotb=3)
print(e_kin.shape, e_mag.shape)
