"""End-to-end golden vectors of the reference for the phase-field branch: samples/phase_field (Chebyshev stage).

The reference's autotest (`samples/phase_field/unitTest.py`, rtol 1e-8) runs 100 CNAB2 steps of Boussinesq convection with a
phase field -- Stefan number 1, tmelt = 0.11, epsPhase = 0.03, penaltyFac = 0.5, Ra = 2e5, Ek = 1e-3, rigid walls, minc = 4,
l_max = 64, n_r_max = 65 -- and compares e_kin.TAG and phase.TAG.  What the radial loop contributes:

  get_nl.f90:333-344   the penalty terms -phi v / (epsPhase penaltyFac)^2 of the advection and the phase-field source phiTerms
  get_td  (rIter.f90:698, updatePhi.f90:173-186)   dphidt = phiTerms on bulk levels
  get_ekin_solid_liquid (rIter.f90:360, outMisc.f90:1169-1221)   ekinSr, ekinLr, volSr -> columns volS, ekinS, ekinL of phase.TAG

Host: oracle/lmloop.py ShellHost(phase=...) (updatePhi, the Stefan coupling of updateS), which also yields the columns rphase,
tphase, fcmb, ficb, dtTPhi (spectral, l = 0); rmelt_mean .. rmelt_max and phase_min / phase_max are taken from the phase field and
temperature on the grid (outMisc.f90:881-899, 952-953).  CPU leg: the oracle's loop for the first rows; GPU leg: the CUDA loop
and magic_rloop_diagnostics(MAGIC_DIAG_PHASE) through the C ABI for all ten logged rows, the oracle's diagnostics beside it.
Fixture: tests/golden/phase_field_reference.npz (tests/golden/make_phase_field_fixture.py).
"""
import os

import numpy as np
import pytest

from magic_b200.riter import DIAG_PHASE

HERE = os.path.dirname(os.path.abspath(__file__))
RTOL = 1e-8                      # samples/phase_field/unitTest.py
OSQ4PI = 1.0 / np.sqrt(4.0 * np.pi)


@pytest.fixture(scope="module")
def golden():
    d = np.load(os.path.join(HERE, "golden", "phase_field_reference.npz"))
    return {k: d[k] for k in d.files}


def _sizes(golden):
    from magic_b200.sht import grid_sizes
    gs = grid_sizes(n_phi_tot=int(golden["n_phi_tot"]), minc=int(golden["minc"]))
    assert (gs["l_max"], gs["n_phi_max"], gs["n_theta_max"]) == (64, 48, 96)
    return gs


def _setup(golden, gs, lm2l, lm2m):
    from magic_b200.workload import make_params
    from oracle.lmloop import ShellHost
    N = int(golden["n_r_max"])
    phase = {k: float(golden[k]) for k in ("stef", "tmelt", "phaseDiffFac", "penaltyFac", "epsPhase")}
    phase.update(ktopphi=int(golden["ktopphi"]), kbotphi=int(golden["kbotphi"]))
    kv = dict(ktopv=int(golden["ktopv"]), kbotv=int(golden["kbotv"]))
    h = ShellHost(lm2l, lm2m, None, n_r_max=N, n_cheb_max=int(golden["n_cheb_max"]), init_s1=int(golden["init_s1"]), l_mag=False,
                  phase=phase, **kv, **{k: float(golden[k]) for k in ("radratio", "ra", "ek", "pr", "dtmax", "alpha", "amp_s1")})
    p = make_params("hydro", N, **kv)
    p.ra, p.CorFac = float(golden["ra"]), 1.0 / float(golden["ek"])
    p.r_cmb, p.r_icb = h.g.r_cmb, h.g.r_icb
    p.courfac, p.alffac = float(golden["courfac"]), float(golden["alffac"])
    p.l_phase_field = 1
    p.epsPhase, p.phaseDiffFac, p.penaltyFac, p.tmelt = (phase[k] for k in ("epsPhase", "phaseDiffFac", "penaltyFac", "tmelt"))
    g, r, one = h.g, h.g.r, np.ones(N)
    delxr2 = np.zeros(N)                                               # preCalculations.f90:304-310
    delxr2[0] = (r[0] - r[1]) ** 2
    delxr2[-1] = (r[-2] - r[-1]) ** 2
    for n in range(1, N - 1):
        delxr2[n] = min(r[n - 1] - r[n], r[n] - r[n + 1]) ** 2
    lR = np.full(N, gs["l_max"])
    rad = dict(nR=np.arange(1, N + 1, dtype=np.int32), l_R=lR.astype(np.int32), r=r, or1=g.or1, or2=g.or2, or4=g.or2 ** 2, orho1=one,
               orho2=one, beta=0 * one, rho0=one, otemp1=one, temp0=one, visc=one, epscProf=one, delxr2=delxr2,
               delxh2=r ** 2 / (lR * (lR + 1.0)))
    rad["lambda"] = one
    return h, p, rad


def lagrange_interp(xp, x, yp):
    """useful.f90 lagrange_interp: the interpolating polynomial through (xp, yp) at x."""
    res = 0.0
    for i in range(len(xp)):
        w = 1.0
        for j in range(len(xp)):
            if j != i:
                w *= (x - xp[j]) / (xp[i] - xp[j])
        res += w * yp[i]
    return res


def rmelt_tmelt(phase, temp, r):
    """get_rmelt_tmelt, outMisc.f90:1319-1376 (0-based levels: CMB = 0, ICB = N - 1)."""
    N = len(r)
    k = 0
    for n in range(1, N):
        if phase[n] < 0.5 and phase[n - 1] > 0.5:
            k = n
    if k == 0:
        return r[0], temp[0]
    if k == 1:
        a, b = k - 1, k + 2
    elif k == N - 1:
        a, b = k - 3, k
    else:
        a, b = k - 2, k + 1
    rph = lagrange_interp(phase[a:b + 1], 0.5, r[a:b + 1])
    return rph, lagrange_interp(r[a:b + 1], rph, temp[a:b + 1])


class PhaseSeries:
    """outPhase, outMisc.f90:826-989: one row of phase.TAG from the per-level sums (slots 32-36) and the host's fields."""

    def __init__(self, h, sht, gauss_rows):
        self.h, self.sht, self.gauss, self.TPhi, self.t_last = h, sht, gauss_rows, 0.0, 0.0

    def row(self, d):
        h, g = self.h, self.h.g
        r, N = g.r, len(g.r)
        lm00 = 0
        assert h.lm2l[lm00] == 0 and h.lm2m[lm00] == 0
        s00, phi00 = h.s[:, lm00].real, h.phi[:, lm00].real
        ds00 = h.ds[:, lm00].real          # get_entropy_rhs_imp's ds of the current s (updateS.f90:683)
        volS, ekinS, ekinL = (g.rInt_R(d[:, c]) for c in (34, 32, 33))
        opr = 1.0 / h.pr
        fcmb = -opr * ds00[0] * OSQ4PI * 4.0 * np.pi * g.r_cmb ** 2
        ficb = -opr * ds00[-1] * OSQ4PI * 4.0 * np.pi * g.r_icb ** 2
        TPhi = 4.0 * np.pi * g.rInt_R(OSQ4PI * (s00 - h.phase["stef"] * phi00) * r * r)
        dtTPhi = (TPhi - self.TPhi) / (h.time - self.t_last) if h.time > self.t_last else 0.0
        self.TPhi, self.t_last = TPhi, h.time
        rphase, tphase = rmelt_tmelt(OSQ4PI * phi00, OSQ4PI * s00, r)
        lmax = int(h.lm2l.max())
        ph = np.stack([self.sht.scal_to_spat(h.phi[n], lmax) for n in range(N)])       # [N, n_phi, n_theta]
        te = np.stack([self.sht.scal_to_spat(h.s[n], lmax) for n in range(N)])
        rm = np.zeros(ph.shape[1:])
        tm = np.zeros(ph.shape[1:])
        for ip in range(ph.shape[1]):
            for it in range(ph.shape[2]):
                rm[ip, it], tm[ip, it] = rmelt_tmelt(ph[:, ip, it], te[:, ip, it], r)
        norm = 0.5 / ph.shape[1]
        return np.array([h.time, rphase, tphase, (self.gauss[None, :] * rm).sum() * norm, (self.gauss[None, :] * tm).sum() * norm, rm.min(), rm.max(),
                         volS, ekinS, ekinL, fcmb, ficb, dtTPhi, d[:, 35].min(), d[:, 36].max()]), (ph.min(), ph.max())


def _check_row(golden, h, series, d, row):
    gk = np.concatenate([[h.time], h.e_kin()])
    np.testing.assert_allclose(gk, golden["e_kin"][row], rtol=RTOL, atol=1e-20, err_msg=f"e_kin row {row}")
    got, (pmin, pmax) = series.row(d)
    assert d[:, 35].min() == pytest.approx(pmin, abs=1e-13) and d[:, 36].max() == pytest.approx(pmax, abs=1e-13)
    if row == 0:
        return                       # phase.TAG skips the first log (outMisc.f90:965); the call primes TPhi as the reference does
    ref = golden["phase"][row - 1]
    np.testing.assert_allclose(got[:12], ref[:12], rtol=RTOL, err_msg=f"phase row {row}")
    # dtTPhi, phase_min, phase_max are printed with six digits (ES13.5); phase_min is the undershoot of the front, O(1e-8)
    np.testing.assert_allclose(got[12:15], ref[12:15], rtol=1e-5, err_msg=f"dtTPhi, phase_min, phase_max row {row}")
    np.testing.assert_allclose(got[8] + got[9], gk[1] + gk[2], rtol=1e-9)   # solid + liquid = the spectral kinetic energy


def _oparams(p):
    from oracle.oracle import Params as OParams
    op = OParams()
    for n, _ in p._fields_:
        setattr(op, n, getattr(p, n))
    return op


def _oracle(gs, golden, fast):
    from oracle.oracle import Oracle
    return Oracle(gs["l_max"], minc=int(golden["minc"]), n_theta=gs["n_theta_max"], n_phi=gs["n_phi_max"], m_max=gs["m_max"],
                  threads=min(4, os.cpu_count() or 1), fast=fast)


def test_oracle_loop_reproduces_phase_field_rows(golden):
    """CPU: the oracle's radial loop (penalty terms, phiTerms, dphidt) and its get_ekin_solid_liquid under the numpy host: rows 0-1 of
    e_kin.TAG and the first row of phase.TAG (MAGIC_PHASE_CPU_ROWS for more; ten steps per row, about a second per step)."""
    gs = _sizes(golden)
    o = _oracle(gs, golden, fast=True)
    h, p, rad = _setup(golden, gs, o.lm2l, o.lm2m)
    op = _oparams(p)
    h.radial_loop = lambda f: o.radial_loop(op, rad, f)
    series = PhaseSeries(h, o, o.gauss)
    diag = lambda: o.radial_diagnostics(op, rad, h.fields_Rloc(), DIAG_PHASE)
    _check_row(golden, h, series, diag(), 0)
    for row in range(1, int(os.environ.get("MAGIC_PHASE_CPU_ROWS", "1")) + 1):
        for _ in range(int(golden["n_log_step"])):
            h.step()
        _check_row(golden, h, series, diag(), row)
    # negative control: without the penalty / phase-field terms in the loop the golden row is missed
    p0 = _oparams(p)
    p0.penaltyFac = 1.0
    h.radial_loop = lambda f: o.radial_loop(p0, rad, f)
    for _ in range(int(golden["n_log_step"])):
        h.step()
    gk = np.concatenate([[h.time], h.e_kin()])
    assert np.abs(gk / golden["e_kin"][row + 1] - 1.0).max() > 1e-4


@pytest.mark.gpu
def test_gpu_loop_reproduces_phase_field_series(golden):
    """The CUDA radial loop with l_phase_field in the time loop and magic_rloop_diagnostics(MAGIC_DIAG_PHASE): all 11 rows of
    e_kin.TAG and all 10 rows of phase.TAG of the Chebyshev stage; the oracle's diagnostics are evaluated on the same fields at
    every row and agree with the device's."""
    from magic_b200 import RadialLoop, Sht
    gs = _sizes(golden)
    s = Sht(gs["l_max"], m_max=gs["m_max"], minc=int(golden["minc"]), n_theta_max=gs["n_theta_max"], n_phi_max=gs["n_phi_max"])
    h, p, rad = _setup(golden, gs, s.lm2l, s.lm2m)
    rl = RadialLoop(s, p, rad)
    h.radial_loop = lambda f: rl.radialLoop(f)
    o = _oracle(gs, golden, fast=False)
    op = _oparams(p)
    series = PhaseSeries(h, s, o.gauss)
    for row in range(len(golden["e_kin"])):
        if row:
            for _ in range(int(golden["n_log_step"])):
                h.step()
        d = rl.diagnostics(h.fields_Rloc(), DIAG_PHASE)
        _check_row(golden, h, series, d, row)
        d_orc = o.radial_diagnostics(op, rad, h.fields_Rloc(), DIAG_PHASE)
        np.testing.assert_allclose(d[:, 32:35], d_orc[:, 32:35], rtol=1e-11, atol=1e-13 * np.abs(d_orc[:, 32:35]).max())
        np.testing.assert_allclose(d[:, 35:37], d_orc[:, 35:37], rtol=0, atol=1e-12)
        assert not d[:, :32].any() and not d[:, 37:].any()
    # the phase sums next to the other diagnostics in one call
    from magic_b200.riter import DIAG_HEL, DIAG_HEMI, DIAG_POWER
    mask = DIAG_PHASE | DIAG_HEL | DIAG_HEMI | DIAG_POWER
    d, d_orc = rl.diagnostics(h.fields_Rloc(), mask), o.radial_diagnostics(op, rad, h.fields_Rloc(), mask)
    scale = np.abs(d_orc).max(axis=0) + 1e-300
    assert (np.abs(d - d_orc).max(axis=0) / scale).max() < 1e-11
    assert np.abs(d_orc[:, [0, 9, 17, 32, 33]]).max(axis=0).min() > 0
    rl.finalize()
    s.finalize_sht()
