"""Torsional-oscillation sums of the radial loop (rIter.f90:395-404): getTOnext's grid part (TO.f90:330-343)
and getTO (TO.f90:141-307).

CPU: the oracle's restatement (oracle/magic_oracle_diag.inc orc_radial_TO) against
  * the same azimuthal means formed in numpy from the oracle's golden-pinned per-call transforms (bulk level),
  * identities that tie it to other routines: sum_theta gauss V2AS = the hemispheric energies of get_hemi, the time-derivative
    terms against BspAS / BpzAS when the kept fields are a scaled copy of the present ones, an axisymmetric flow has no Reynolds
    stress, a rigid rotation gives VAS = Omega r sin(theta) and a Coriolis term of zero.
GPU (-m gpu): magic_rloop_to_next / magic_rloop_to through the C ABI against the oracle on the same seeded spectra (MHD with
rigid rotating walls, stress-free walls, a phase field; more levels than one chunk; device pointers).
"parity unpinned": samples/testTOGeosOutputs compares TO movie points and TOnhs / TOshs columns that only exist after outTO's
cylindrical integration, which the host restatement does not cover.
"""
import numpy as np
import pytest

from magic_b200.riter import DIAG_HEMI, DIAG_RMSBULK, NTO
from magic_b200.workload import make_fields, make_params, make_radial

NAMES = ["V2AS", "VAS", "dzCorAS", "dzRstrAS", "dzAstrAS", "dzLFAS", "Bs2AS", "BspAS", "BpzAS", "BszAS", "BspdAS", "BpsdAS", "BzpdAS", "BpzdAS",
         "dzPenAS"]


def _oracle(l_max, minc=1):
    from oracle.oracle import Oracle
    return Oracle(l_max, minc=minc)


def _oparams(p):
    from oracle.oracle import Params as OParams
    op = OParams()
    for n, _ in p._fields_:
        setattr(op, n, getattr(p, n))
    return op


def _ordered(o):
    """index of every row of the reference's scrambled colatitude layout in geographic order (n_theta_cal2ord)"""
    n = o.n_theta
    return np.array([t // 2 if t % 2 == 0 else n - 1 - t // 2 for t in range(n)])


def _case(physics, l_max, n_r_max, lm2l, lm2m, seed, phase=False, **kw):
    p = make_params(physics, n_r_max, **kw)
    rad = make_radial(n_r_max, l_max)
    f = make_fields(physics, lm2l, lm2m, n_r_max, seed)
    if phase:
        p.l_phase_field, p.epsPhase, p.penaltyFac, p.phaseDiffFac, p.tmelt = 1, 0.05, 0.7, 1.0, 0.3
        f["phi"] = 0.4 * f["s"] + 0.2 * f["w"]
    return p, rad, f


def test_oracle_getTO_against_means_of_the_per_call_transforms():
    l_max, n_r = 16, 4
    o = _oracle(l_max)
    p, rad, f = _case("mhd", l_max, n_r, o.lm2l, o.lm2m, 11)
    p.CorFac = 1234.5
    op = _oparams(p)
    dt = 3e-4
    f_old = {k: (0.9 * v if k in ("b", "db", "ddb", "aj", "dj") else v) for k, v in f.items()}
    last = o.radial_TO(op, rad, f_old, 0)
    got = o.radial_TO(op, rad, f, 1, dtLast=dt, last=last)
    ordr = _ordered(o)
    st, ct = np.sin(np.arccos(o.cosTheta)), o.cosTheta       # scrambled layout, per row
    for i in (1, 2):                                          # bulk levels
        or1, or2 = rad["or1"][i], rad["or2"][i]
        or3, or4 = or1 * or2, or2 * or2
        vr, vt, vp = o.torpol_to_spat(f["w"][i], f["dw"][i], f["z"][i], l_max)          # [n_phi, n_theta]
        _, _, dvpdr = o.torpol_to_spat(f["dw"][i], f["ddw"][i], f["dz"][i], l_max)
        cvr = o.pol_to_curlr_spat(f["z"][i], l_max)
        br, bt, bp = o.torpol_to_spat(f["b"][i], f["db"][i], f["aj"][i], l_max)
        cbr, cbt, _ = o.torpol_to_curl_spat(or2, f["b"][i], f["ddb"][i], f["aj"][i], f["dj"][i], l_max)
        mean = lambda x: x.mean(axis=0)
        ref = {}
        ref["V2AS"] = mean(or4 * vr ** 2 + or2 / st ** 2 * (vt ** 2 + vp ** 2))
        ref["VAS"] = or1 / st * mean(vp)
        ref["dzCorAS"] = -2.0 * p.CorFac * (or2 * st * mean(vr) + or1 * ct / st * mean(vt))
        stress = -or3 / st * (mean(vr * dvpdr) + mean(vt * cvr))                          # Reynolds + axisymmetric part
        ref["dzAstrAS"] = -or3 / st * (mean(vr) * mean(dvpdr) + mean(vt) * mean(cvr))
        ref["dzRstrAS"] = stress - ref["dzAstrAS"]
        ref["dzLFAS"] = or3 / st * mean(cbr * bt - cbt * br)
        bs, bpl, bz = st * or2 * br + ct / st * or1 * bt, or1 * bp / st, ct * or2 * br - or1 * bt
        ref["Bs2AS"], ref["BspAS"], ref["BpzAS"], ref["BszAS"] = mean(bs * bs), mean(bs * bpl), mean(bpl * bz), mean(bs * bz)
        ref["BspdAS"] = ref["BpsdAS"] = 0.1 * ref["BspAS"] / dt                           # kept fields = 0.9 x the present ones
        ref["BzpdAS"] = ref["BpzdAS"] = 0.1 * ref["BpzAS"] / dt
        for nm, r in ref.items():
            g = got[i, NAMES.index(nm)][ordr]
            assert np.abs(g - r).max() < 1e-12 * np.abs(r).max(), nm
        assert not got[i, NAMES.index("dzPenAS")].any()


def test_oracle_getTO_identities():
    l_max, n_r = 16, 5
    o = _oracle(l_max)
    p, rad, f = _case("mhd", l_max, n_r, o.lm2l, o.lm2m, 5)
    op = _oparams(p)
    got = o.radial_TO(op, rad, f, 1, dtLast=1.0)
    # sum_theta gauss V2AS = (1 / 2 pi) int u^2 dOmega = (2 / 2 pi) or2 (ekin_N + ekin_S) of get_hemi (orho1 = 1; boundaries as bulk)
    gauss_ord = np.zeros(o.n_theta)
    gauss_ord[_ordered(o)] = o.gauss
    d = o.radial_diagnostics(op, rad, f, DIAG_HEMI | DIAG_RMSBULK)
    p2 = _oparams(p)
    p2.ktopv = p2.kbotv = 0  # not a boundary type: every level is treated as bulk, as lRmsCalc does for get_hemi above
    got_bulk = o.radial_TO(p2, rad, f, 1, dtLast=1.0)
    np.testing.assert_allclose((got_bulk[:, 0] * gauss_ord).sum(axis=1), rad["or2"] * (d[:, 9] + d[:, 10]) / np.pi, rtol=1e-12)
    # nothing kept (last = 0): the "time derivatives" are the products themselves
    np.testing.assert_allclose(got[:, NAMES.index("BspdAS")], got[:, NAMES.index("BspAS")], rtol=1e-12, atol=1e-12 * np.abs(got[:, 7]).max())
    np.testing.assert_allclose(got[:, NAMES.index("BpzdAS")], got[:, NAMES.index("BpzAS")], rtol=1e-12, atol=1e-12 * np.abs(got[:, 8]).max())
    # rigid walls: no flow sums on the boundary levels except the wall rotation (omega = 0 here)
    assert not got[0, :5].any() and not got[-1, :5].any()
    # an axisymmetric flow has no Reynolds stress
    fa = {k: np.where(o.lm2m[None, :] == 0, v, 0.0) for k, v in f.items()}
    ga = o.radial_TO(op, rad, fa, 1, dtLast=1.0)
    assert np.abs(ga[:, 3]).max() < 1e-12 * np.abs(ga[:, 4]).max()
    # rigid rotation u_phi = Omega r sin(theta): z_10 = c r^2 with Omega = c sqrt(3 / 4 pi); VAS = u_phi, no Coriolis torque term
    lm10 = int(np.where((o.lm2l == 1) & (o.lm2m == 0))[0][0])
    fr = {k: np.zeros_like(v) for k, v in f.items()}
    c = 0.7
    fr["z"][:, lm10] = c * rad["r"] ** 2
    fr["dz"][:, lm10] = 2 * c * rad["r"]
    gr = o.radial_TO(p2, rad, fr, 1, dtLast=1.0)
    sin_ord = np.zeros(o.n_theta)
    sin_ord[_ordered(o)] = np.sin(np.arccos(o.cosTheta))
    np.testing.assert_allclose(gr[:, 1], c * np.sqrt(3.0 / (4.0 * np.pi)) * rad["r"][:, None] * sin_ord[None, :], rtol=1e-12)
    assert np.abs(gr[:, 2]).max() < 1e-12 * p.CorFac * c


def _compare(got, ref, label, tol=1e-12):
    worst = 0.0
    for q in range(NTO):
        scale = np.abs(ref[:, q]).max()
        if scale == 0.0:
            assert np.abs(got[:, q]).max() == 0.0, f"{label}: {NAMES[q]} should be zero"
            continue
        err = np.abs(got[:, q] - ref[:, q]).max() / scale
        worst = max(worst, err)
        assert err < tol, f"{label}: {NAMES[q]} deviates by {err:.2e}"
    print(f"{label}: worst array error {worst:.2e}")


@pytest.mark.gpu
@pytest.mark.parametrize("physics,l_max,n_r,ktopv,kbotv,omega_ic,phase", [("mhd", 21, 7, 2, 2, 2.5, False), ("mhd", 16, 6, 1, 1, 0.0, False),
                                                                          ("mhd", 16, 5, 2, 1, 0.0, True), ("hydro", 32, 40, 2, 2, 0.0, True)])
def test_gpu_to_sums_match_oracle(physics, l_max, n_r, ktopv, kbotv, omega_ic, phase):
    from magic_b200 import RadialLoop, Sht
    s = Sht(l_max)
    o = _oracle(l_max)
    p, rad, f = _case(physics, l_max, n_r, s.lm2l, s.lm2m, 23, phase=phase, ktopv=ktopv, kbotv=kbotv)
    p.omega_ic, p.CorFac = omega_ic, 777.0
    op = _oparams(p)
    f_old = {k: (0.8 * v + 0.1 * f["w"] if k in ("b", "db", "ddb", "aj", "dj") else v) for k, v in f.items()}
    dt = 2.5e-4
    ref = o.radial_TO(op, rad, f, 1, dtLast=dt, last=o.radial_TO(op, rad, f_old, 0))
    rl = RadialLoop(s, p, rad)
    rl.to_next(f_old)
    got = rl.to(f, dt)
    _compare(got, ref, f"TO {physics} l{l_max} ktopv{ktopv} kbotv{kbotv}")
    again = rl.to(f, dt)
    assert np.array_equal(got, again)                      # fixed-shape reductions
    rl.finalize()
    s.finalize_sht()


@pytest.mark.gpu
def test_gpu_to_with_device_pointers():
    import torch
    from magic_b200 import RadialLoop, Sht
    l_max, n_r = 16, 6
    s = Sht(l_max)
    p, rad, f = _case("mhd", l_max, n_r, s.lm2l, s.lm2m, 4)
    rl = RadialLoop(s, p, rad)
    rl.to_next(f)
    ref = rl.to(f, 1e-3)
    keep = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in f.items()}
    ptrs = {k: v.data_ptr() for k, v in keep.items()}
    torch.cuda.synchronize()
    rl.to_next(ptrs, device=True)
    got = rl.to(ptrs, 1e-3, device=True)
    assert np.array_equal(got, ref)
    rl.finalize()
    s.finalize_sht()
