"""End-to-end golden vectors of the reference for the in-loop diagnostics: samples/testOutputs.

The reference's autotest (`samples/testOutputs/unitTest.py`, rtol 1e-8) runs 100 CNAB2 steps of a weakly stratified anelastic
dynamo (strat = 0.1, polytropic index 2, Ra = 3e5, Ek = 1e-3, Pm = 5, rigid insulating walls, l_max = 85, n_r_max = 73 with
n_cheb_max = 71) with l_hel, l_hemi, l_power and l_RMS on, and compares the concatenation of ten time series.  Three of them
are radial integrals of what the radial loop sums on the grid at log steps (rIter.f90:320-342):

  helicity.TAG  (outMisc.f90:329-414)  from get_helicity's HelASr, Hel2ASr, HelnaASr, Helna2ASr      -- 8 columns
  hemi.TAG      (outMisc.f90:243-327)  from get_hemi's hemi_ekin_r, hemi_emag_r (the asymmetry columns round to 0: the run
                                        stays equatorially symmetric)                                  -- 2 columns
  power.TAG     (power.f90:196-353)    column 6 = -int viscASr dr from get_visc_heat                   -- 1 column

With l_RMS on, lRmsCalc treats the boundary levels as bulk on log steps (rIter.f90:215): MAGIC_DIAG_RMSBULK.
Host: oracle/lmloop.py ShellHost, which reproduces e_kin.TAG / e_mag_oc.TAG of this run as well (checked here first).  The
diagnostics come from magic_rloop_diagnostics through the C ABI with the CUDA radial loop in the time loop (GPU test: all 11 rows)
and from the CPU oracle on the same fields (rows 1 and 5 there; row 0 in the CPU test, more with MAGIC_TESTOUTPUTS_CPU_ROWS).  tests/golden/testOutputs_reference.npz holds the five series
(tests/golden/make_testOutputs_fixture.py).
"""
import os

import numpy as np
import pytest

from magic_b200.riter import DIAG_HEL, DIAG_HEMI, DIAG_POWER, DIAG_RMSBULK

HERE = os.path.dirname(os.path.abspath(__file__))
RTOL, ATOL = 1e-8, 1e-20          # samples/testOutputs/unitTest.py
MASK = DIAG_HEL | DIAG_HEMI | DIAG_POWER | DIAG_RMSBULK


@pytest.fixture(scope="module")
def golden():
    d = np.load(os.path.join(HERE, "golden", "testOutputs_reference.npz"))
    return {k: d[k] for k in d.files}


def _sizes(golden):
    from magic_b200.sht import grid_sizes
    gs = grid_sizes(n_phi_tot=int(golden["n_phi_tot"]))
    assert (gs["l_max"], gs["lm_max"]) == (85, 3741)
    return gs


def _setup(golden, lm2l, lm2m):
    from magic_b200.workload import make_params
    from oracle.lmloop import ShellHost
    N = int(golden["n_r_max"])
    kw = {k: float(golden[k]) for k in ("radratio", "ra", "ek", "pr", "prmag", "dtmax", "alpha", "amp_s1", "amp_b1", "strat", "polind",
                                        "g0", "g1", "g2")}
    h = ShellHost(lm2l, lm2m, None, n_r_max=N, n_cheb_max=int(golden["n_cheb_max"]), init_s1=int(golden["init_s1"]),
                  init_b1=int(golden["init_b1"]), l_mag=True, ktopv=int(golden["ktopv"]), kbotv=int(golden["kbotv"]), **kw)
    p = make_params("anel", N, ktopv=int(golden["ktopv"]), kbotv=int(golden["kbotv"]))
    p.l_mag = p.l_mag_nl = p.l_mag_LF = 1
    p.ViscHeatFac, p.OhmLossFac = h.ViscHeatFac, h.OhmLossFac          # radial.f90:762-764
    p.ra, p.CorFac, p.LFfac, p.opm = kw["ra"], 1.0 / kw["ek"], h.LFfac, h.opm
    p.r_cmb, p.r_icb = h.g.r_cmb, h.g.r_icb
    p.courfac, p.alffac = float(golden["courfac"]), float(golden["alffac"])
    g, r, one = h.g, h.g.r, np.ones(N)
    delxr2 = np.zeros(N)                                               # preCalculations.f90:304-310
    delxr2[0] = (r[0] - r[1]) ** 2
    delxr2[-1] = (r[-2] - r[-1]) ** 2
    for n in range(1, N - 1):
        delxr2[n] = min(r[n - 1] - r[n], r[n] - r[n + 1]) ** 2
    lR = np.full(N, 85)
    rad = dict(nR=np.arange(1, N + 1, dtype=np.int32), l_R=lR.astype(np.int32), r=r, or1=g.or1, or2=g.or2, or4=g.or2 ** 2,
               orho1=1.0 / h.rho0, orho2=1.0 / h.rho0 ** 2, beta=h.beta, rho0=h.rho0, otemp1=1.0 / h.temp0, temp0=h.temp0,
               visc=one, epscProf=one, delxr2=delxr2, delxh2=r ** 2 / (lR * (lR + 1.0)))
    rad["lambda"] = h.lam
    return h, p, rad


def _round_off(x, ref=1.0, fac=1e3):
    """useful.f90:304-330"""
    return 0.0 if abs(x) < fac * np.finfo(float).eps * abs(ref) else x


def helicity_row(h, d):
    """outHelicity, outMisc.f90:361-408, from the per-level sums of get_helicity (slots 0-8)."""
    g = h.g
    r2 = g.r ** 2
    vol_oc = 4.0 / 3.0 * np.pi * (g.r_cmb ** 3 - g.r_icb ** 3)
    I = lambda col: g.rInt_R(d[:, col] * r2)
    HelN, HelS, HelnaN, HelnaS = (2 * np.pi * I(c) / (vol_oc / 2) for c in (0, 1, 4, 5))
    HelRMSN, HelRMSS, HelnaRMSN, HelnaRMSS = (np.sqrt(2 * np.pi * I(c) / (vol_oc / 2)) for c in (2, 3, 6, 7))
    if HelnaRMSN + HelnaRMSS != 0:
        HelnaN, HelnaS = HelnaN / HelnaRMSN, HelnaS / HelnaRMSS
    else:
        HelnaN = HelnaS = 0.0
    if HelRMSN + HelRMSS != 0:
        HelN, HelS = HelN / HelRMSN, HelS / HelRMSS
    else:
        HelN = HelS = 0.0
    return np.array([h.time, HelN, HelS, HelRMSN, HelRMSS, HelnaN, HelnaS, HelnaRMSN, HelnaRMSS])


def hemi_row(h, d):
    """outHemi, outMisc.f90:276-321, from the per-level sums of get_hemi (slots 9-16); eScale = vScale = 1."""
    g = h.g
    I = lambda col: g.rInt_R(d[:, col])
    ekinN, ekinS, vrabsN, vrabsS = I(9), I(10), I(11), I(12)
    emagN, emagS, brabsN, brabsS = h.LFfac * I(13), h.LFfac * I(14), I(15), I(16)
    hemi_emag = hemi_br = hemi_cmb = hemi_ekin = hemi_vr = 0.0
    if emagN + emagS > 0:
        hemi_emag = abs(emagN - emagS) / (emagN + emagS)
        hemi_br = abs(brabsN - brabsS) / (brabsN + brabsS)
        hemi_cmb = abs(d[0, 15] - d[0, 16]) / (d[0, 15] + d[0, 16])
    if ekinN + ekinS > 0:
        hemi_ekin = abs(ekinN - ekinS) / (ekinN + ekinS)
        hemi_vr = abs(vrabsN - vrabsS) / (vrabsN + vrabsS)
    return np.array([h.time] + [_round_off(x) for x in (hemi_vr, hemi_ekin, hemi_br, hemi_emag, hemi_cmb)] + [ekinN + ekinS, emagN + emagS])


def visc_diss(h, d):
    """power.f90:214,289: viscDiss = -eScale rInt_R(viscASr)."""
    return -h.g.rInt_R(d[:, 17])


def _check(golden, h, diag, row):
    gk = np.concatenate([[h.time], h.e_kin()])
    gm = np.concatenate([[h.time], h.e_mag_oc()])
    np.testing.assert_allclose(gk, golden["e_kin"][row], rtol=RTOL, atol=ATOL, err_msg=f"e_kin row {row}")
    np.testing.assert_allclose(gm, golden["e_mag_oc"][row], rtol=RTOL, atol=ATOL, err_msg=f"e_mag_oc row {row}")
    d = diag(h.fields_Rloc())
    np.testing.assert_allclose(helicity_row(h, d), golden["helicity"][row], rtol=RTOL, atol=ATOL, err_msg=f"helicity row {row}")
    np.testing.assert_allclose(hemi_row(h, d), golden["hemi"][row], rtol=RTOL, atol=ATOL, err_msg=f"hemi row {row}")
    if row >= 1:   # power.TAG starts at the second log step (power.f90:343)
        pw = golden["power"][row - 1]
        assert abs(pw[0] - h.time) < 1e-12
        np.testing.assert_allclose(visc_diss(h, d), pw[5], rtol=RTOL, err_msg=f"power (viscDiss) row {row}")
    return d


def _run(golden, h, diag, n_rows):
    for row in range(1, n_rows + 1):
        for _ in range(int(golden["n_log_step"])):
            h.step()
        _check(golden, h, diag, row)


def _oparams(p):
    from oracle.oracle import Params as OParams
    op = OParams()
    for n, _ in p._fields_:
        setattr(op, n, getattr(p, n))
    return op


def _oracle(golden, fast):
    from oracle.oracle import Oracle
    gs = _sizes(golden)
    return Oracle(gs["l_max"], n_theta=gs["n_theta_max"], n_phi=gs["n_phi_max"], m_max=gs["m_max"], threads=min(4, os.cpu_count() or 1),
                  fast=fast)


def _negative_controls(golden, h, o, op, rad, row):
    """With the non-axisymmetric helicity fed the full fields, or a viscous heating without the density-gradient (beta) terms,
    the golden row is missed (the beta terms cancel in Hel itself)."""
    d = o.radial_diagnostics(op, rad, h.fields_Rloc(), MASK)
    ref = golden["helicity"][row]
    swapped = d.copy()
    swapped[:, 4:8] = d[:, 0:4]
    assert np.abs(helicity_row(h, swapped)[5:] / ref[5:] - 1.0).max() > 1e-3
    d2 = o.radial_diagnostics(op, dict(rad, beta=0 * rad["beta"]), h.fields_Rloc(), MASK)
    assert abs(visc_diss(h, d2) / golden["power"][row - 1][5] - 1.0) > 1e-4


def test_oracle_diagnostics_on_the_start_fields(golden):
    """CPU: row 0 (start fields: no flow, the imposed field) of e_kin, e_mag_oc, helicity and hemi with the oracle's diagnostics.
    The later rows need ten time steps each at l_max = 85 (about a minute per row on four host cores), so the oracle's diagnostics
    are pinned to them in the GPU leg below, where the CUDA loop does the stepping; MAGIC_TESTOUTPUTS_CPU_ROWS=1 runs the first
    logged row here as well."""
    o = _oracle(golden, fast=True)
    h, p, rad = _setup(golden, o.lm2l, o.lm2m)
    op = _oparams(p)
    h.radial_loop = lambda f: o.radial_loop(op, rad, f)
    diag = lambda f: o.radial_diagnostics(op, rad, f, MASK)
    _check(golden, h, diag, 0)
    rows = int(os.environ.get("MAGIC_TESTOUTPUTS_CPU_ROWS", "0"))
    if rows:
        _run(golden, h, diag, rows)
        _negative_controls(golden, h, o, op, rad, rows)


@pytest.mark.gpu
def test_gpu_diagnostics_reproduce_helicity_hemi_and_power(golden):
    """magic_rloop_diagnostics (host field pointers, what rIter_cuda_t holds) with the CUDA radial loop in the time loop: all 11
    logged rows (100 steps).  On rows 1 and 5 the CPU oracle's diagnostics are evaluated on the same fields and held against the
    same golden rows (the pin of oracle/magic_oracle_diag.inc), followed by the negative controls."""
    from magic_b200 import RadialLoop, Sht
    gs = _sizes(golden)
    s = Sht(gs["l_max"], m_max=gs["m_max"], n_theta_max=gs["n_theta_max"], n_phi_max=gs["n_phi_max"])
    h, p, rad = _setup(golden, s.lm2l, s.lm2m)
    rl = RadialLoop(s, p, rad)
    h.radial_loop = lambda f: rl.radialLoop(f)
    diag = lambda f: rl.diagnostics(f, MASK)
    o = _oracle(golden, fast=False)
    op = _oparams(p)
    _check(golden, h, diag, 0)
    for row in range(1, len(golden["e_kin"])):
        for _ in range(int(golden["n_log_step"])):
            h.step()
        d = _check(golden, h, diag, row)
        if row in (1, 5):
            d_orc = _check(golden, h, lambda f: o.radial_diagnostics(op, rad, f, MASK), row)
            scale = np.abs(d_orc).max(axis=0) + 1e-300
            assert (np.abs(d - d_orc).max(axis=0) / scale).max() < 1e-12
    _negative_controls(golden, h, o, op, rad, len(golden["e_kin"]) - 1)
    rl.finalize()
    s.finalize_sht()
