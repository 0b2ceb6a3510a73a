"""Extracts the golden vectors of samples/phase_field from the reference tree (run in the build container).

samples/phase_field/unitTest.py compares, at rtol 1e-8, `cat e_kin.start e_kin.continue` with reference.out and
`cat phase.start phase.continue` with referencePhase.out.  The first run (input_cheb.nml, tag "start") is a Chebyshev run
from scratch: Boussinesq convection with a phase field (Stefan number 1, tmelt = 0.11, epsPhase = 0.03, penaltyFac = 0.5,
phaseDiffFac = 1, ktopphi = kbotphi = 2), Ra = 2e5, Ek = 1e-3, Pr = 1, rigid walls, n_phi_tot = 192 with minc = 4
-> l_max = 64, n_r_max = 65 with n_cheb_max = 63, CNAB2 with dt = 1e-4 from init_s1 = 404 (amp 0.1), 100 steps logged every 10:
rows 0-10 of reference.out and rows 0-9 of referencePhase.out (phase.TAG skips the first log, outMisc.f90:965).  The second run
(input_FD.nml) restarts from the checkpoint with finite differences and is outside what the host restatement covers.
phase.TAG columns (outMisc.f90:971-974): time, rphase, tphase, rmelt_mean, tmelt_mean, rmelt_min, rmelt_max, volS, ekinS, ekinL,
fcmb, ficb, dtTPhi, phase_min, phase_max; volS / ekinS / ekinL are radial integrals of what get_ekin_solid_liquid sums on the grid
inside the radial loop (rIter.f90:360, outMisc.f90:1169-1221).
"""
import os
import re

import numpy as np

REF = "/root/reference/samples/phase_field"
HERE = os.path.dirname(os.path.abspath(__file__))

nml = open(os.path.join(REF, "input_cheb.nml")).read()


def val(name, cast=float):
    m = re.search(r"^\s*" + name + r"\s*=\s*([^,\s]+)", nml, re.M | re.I)
    assert m, name
    return cast(m.group(1).replace("D", "e").replace("d", "e"))


e_kin = np.loadtxt(os.path.join(REF, "reference.out"))
phase = np.loadtxt(os.path.join(REF, "referencePhase.out"))
assert e_kin.shape == (17, 9) and phase.shape == (15, 15)
n_steps, n_log = val("n_time_steps", int), val("n_log_step", int)
n_rows = n_steps // n_log
par = {k: val(k) for k in ("ra", "ek", "pr", "radratio", "stef", "tmelt", "phaseDiffFac", "penaltyFac", "epsPhase", "dtmax", "alpha", "amp_s1",
                           "courfac", "alffac")}
ipar = {k: val(k, int) for k in ("n_r_max", "n_cheb_max", "n_phi_tot", "minc", "init_s1", "ktops", "kbots", "ktopv", "kbotv", "ktopphi",
                                 "kbotphi")}
print(par, ipar, n_rows)
np.savez_compressed(os.path.join(HERE, "phase_field_reference.npz"), n_log_step=n_log, e_kin=e_kin[:n_rows + 1], phase=phase[:n_rows],
                    **par, **ipar)
