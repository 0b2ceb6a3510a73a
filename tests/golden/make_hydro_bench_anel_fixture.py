"""Extracts the golden vectors of samples/hydro_bench_anel from the reference tree (run in the build container).

reference.out is the e_kin.TAG series MagIC's autotest compares against at rtol 1e-8: 300 CNAB2 steps of dt=1e-4 of the
anelastic hydro benchmark (strat=5, polind=2, g2=1, stress-free walls, l_correct_AMz/AMe, n_phi_tot=288 -> l_max=96,
n_r_max=97, n_cheb_max=95), logged every 10 steps (31 rows).  The values of input.nml the host restatement needs are
stored next to it.
"""
import os

import numpy as np

REF = "/root/reference/samples/hydro_bench_anel"
HERE = os.path.dirname(os.path.abspath(__file__))

e_kin = np.loadtxt(os.path.join(REF, "reference.out"))
np.savez_compressed(os.path.join(HERE, "hydro_bench_anel_reference.npz"), e_kin=e_kin, n_log_step=10,
                    n_r_max=97, n_cheb_max=95, n_phi_tot=288, minc=1, ra=1.48638035e5, ek=1e-3, pr=1.0, prmag=5.0,
                    strat=5.0, polind=2.0, radratio=0.35, g0=0.0, g1=0.0, g2=1.0, dtmax=1e-4, alpha=0.6, init_s1=1919,
                    amp_s1=0.01, ktopv=1, kbotv=1, courfac=2.5, alffac=1.0)
print(e_kin.shape)
