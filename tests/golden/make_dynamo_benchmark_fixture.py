"""Extracts the golden vectors of samples/dynamo_benchmark from the reference tree (run in the build container).

reference.out / referenceMag.out are the e_kin.TAG / e_mag_oc.TAG series MagIC's autotest compares against at
rtol 1e-8 (samples/dynamo_benchmark/unitTest.py:97-106); rows 0..200 (t=0 and the first 200 CNAB2 steps of dt=1e-4) are
kept, with the values of input.nml the host restatement needs.
"""
import os

import numpy as np

REF = "/root/reference/samples/dynamo_benchmark"
HERE = os.path.dirname(os.path.abspath(__file__))

e_kin = np.loadtxt(os.path.join(REF, "reference.out"))[:201]
e_mag = np.loadtxt(os.path.join(REF, "referenceMag.out"))[:201]
np.savez_compressed(os.path.join(HERE, "dynamo_benchmark_reference.npz"), e_kin=e_kin, e_mag_oc=e_mag,
                    n_r_max=33, l_max=16, minc=1, ra=1e5, ek=1e-3, pr=1.0, prmag=5.0, radratio=0.35, dtmax=1e-4,
                    alpha=0.6, init_s1=404, amp_s1=0.1, init_b1=3, amp_b1=5.0, courfac=2.5, alffac=1.0)
print(e_kin.shape, e_mag.shape)
