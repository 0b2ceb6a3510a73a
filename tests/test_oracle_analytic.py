"""Pins the CPU oracle with analytic known answers and an independent scipy evaluation.

The reference holds no transform-level golden vectors and cannot be compiled here; these tests anchor
each transform on its own, next to the two golden-vector tests of reference outputs (tests/test_reference_energy.py,
tests/test_dynamo_benchmark.py).
"""
import numpy as np
import pytest
from scipy.special import gammaln, lpmv

from oracle.oracle import Oracle, get_blocks, grid_sizes
from tests.util import random_spectrum, rel_l2


@pytest.fixture(scope="module")
def o16():
    return Oracle(16)


@pytest.fixture(scope="module")
def o32m3():
    gs = grid_sizes(n_phi_tot=96, minc=3)
    return Oracle(gs["l_max"], minc=3, n_theta=gs["n_theta_max"], n_phi=gs["n_phi_max"], m_max=gs["m_max"])


def test_grid_sizes_match_survey_table():
    # SURVEY.md section 8 table (truncation.f90:55-105, nalias=20)
    assert grid_sizes(l_max=16) == dict(l_max=16, m_max=16, n_theta_max=24, n_phi_max=48, n_m_max=17, lm_max=153,
                                        n_phi_tot=48)
    g = grid_sizes(n_phi_tot=288)
    assert (g["l_max"], g["n_theta_max"], g["lm_max"]) == (96, 144, 4753)
    g = grid_sizes(l_max=255)
    assert (g["n_theta_max"], g["n_phi_max"], g["lm_max"]) == (384, 768, 32896)
    g = grid_sizes(l_max=511)
    assert (g["n_theta_max"], g["n_phi_max"], g["lm_max"]) == (768, 1536, 131328)
    g = grid_sizes(l_max=1023)
    assert (g["n_theta_max"], g["n_phi_max"], g["lm_max"]) == (1536, 3072, 524800)
    g = grid_sizes(n_phi_tot=96, minc=3)
    assert (g["l_max"], g["m_max"], g["n_theta_max"], g["n_phi_max"], g["n_m_max"], g["lm_max"]) == (32, 30, 48, 32, 11, 198)


def test_get_blocks_remainder_goes_to_last_ranks():
    s, e = get_blocks(257, 8)  # parallel.f90:75-92
    assert list(e - s + 1) == [32] * 7 + [33]
    s, e = get_blocks(121, 4)
    assert list(e - s + 1) == [30, 30, 30, 31]
    assert s[0] == 1 and e[-1] == 121


def test_gauss_nodes_against_numpy(o16):
    x, w = np.polynomial.legendre.leggauss(o16.n_theta)
    assert np.allclose(np.cos(o16.theta_ord), x[::-1], atol=1e-15)
    # scrambled weights: rows 2k, 2k+1 carry the same weight
    assert np.allclose(o16.gauss[0::2], w[::-1][: o16.n_theta // 2], atol=1e-15)
    assert np.allclose(o16.gauss[0::2], o16.gauss[1::2])


@pytest.mark.parametrize("ctx", ["o16", "o32m3"])
def test_plm_table_against_scipy(ctx, request):
    """plms.f90 norm=2: orthonormal, no Condon-Shortley phase."""
    o = request.getfixturevalue(ctx)
    P = o.plm()
    x = np.cos(o.theta_ord[: o.n_theta // 2])
    for lm in range(o.lm_max):
        l, m = int(o.lm2l[lm]), int(o.lm2m[lm])
        norm = np.exp(0.5 * (np.log((2 * l + 1) / (4 * np.pi)) + gammaln(l - m + 1) - gammaln(l + m + 1)))
        ref = (-1) ** m * norm * lpmv(m, l, x)
        assert np.allclose(P[:, lm], ref, rtol=1e-11, atol=1e-13), (l, m)


def test_dplm_is_sin_dtheta_plm(o16):
    """dPlm = sin(theta) dP/dtheta, checked with the unnormalised identity
    sin(theta) dP_l^m/dtheta = l cos(theta) P_l^m - (l+m) P_{l-1}^m."""
    o = o16
    D = o.dplm()
    th = o.theta_ord[: o.n_theta // 2]
    x = np.cos(th)
    for lm in range(o.lm_max):
        l, m = int(o.lm2l[lm]), int(o.lm2m[lm])
        norm = np.exp(0.5 * (np.log((2 * l + 1) / (4 * np.pi)) + gammaln(l - m + 1) - gammaln(l + m + 1)))
        pl = lpmv(m, l, x)
        plm1 = lpmv(m, l - 1, x) if l - 1 >= m else 0.0
        ref = (-1) ** m * norm * (l * x * pl - (l + m) * plm1)
        assert np.allclose(D[:, lm], ref, rtol=1e-10, atol=1e-12), (l, m)


def test_known_harmonics(o16):
    o = o16
    S = np.zeros(o.lm_max, complex)
    S[0] = 1
    assert np.allclose(o.scal_to_spat(S, 16), 1 / np.sqrt(4 * np.pi), atol=1e-15)
    S[:] = 0
    S[1] = 1  # Y10 with N/S sign flip on interleaved rows
    assert np.allclose(o.scal_to_spat(S, 16), np.sqrt(3 / 4 / np.pi) * o.cosTheta[None, :], atol=1e-15)
    lm11 = int(np.where((o.lm2l == 1) & (o.lm2m == 1))[0][0])
    S[:] = 0
    S[lm11] = 1  # no CS phase, factor 2 from Hermitian completion
    phi = 2 * np.pi * np.arange(o.n_phi) / o.n_phi
    assert np.allclose(o.scal_to_spat(S, 16), 2 * np.sqrt(3 / 8 / np.pi) * o.sinTheta[None, :] * np.cos(phi)[:, None],
                       atol=1e-14)


def test_fft_against_numpy(o16):
    o = o16
    rng = np.random.default_rng(1)
    F = rng.standard_normal((o.n_phi // 2 + 1, o.n_theta)) + 1j * rng.standard_normal((o.n_phi // 2 + 1, o.n_theta))
    g = o.ifft_many(F)
    assert np.allclose(g, np.fft.irfft(F, n=o.n_phi, axis=0) * o.n_phi, atol=1e-12)
    assert np.allclose(o.fft_many(g), np.fft.rfft(g, axis=0) / o.n_phi, atol=1e-13)


@pytest.mark.parametrize("ctx,lcut", [("o16", 16), ("o16", 11), ("o32m3", 32), ("o32m3", 20)])
def test_roundtrips_parseval_and_lcut(ctx, lcut, request):
    o = request.getfixturevalue(ctx)
    rng = np.random.default_rng(7)
    S = random_spectrum(o, rng)
    T = random_spectrum(o, rng, zero_l0=True)
    keep = o.lm2l <= lcut
    f = o.scal_to_spat(S, lcut)
    S2 = o.scal_to_SH(f, lcut)
    assert rel_l2(S2[keep], S[keep]) < 1e-13
    assert np.all(S2[~keep] == 0)
    # Parseval: int f^2 dOmega = sum (2 - delta_m0) |S_lm|^2   (over the minc-fold sector x minc)
    w = 2 * np.pi / (o.n_phi) * o.gauss  # dphi over the full circle divided by points
    lhs = np.sum(f ** 2 * w[None, :])
    rhs = np.sum(np.where(o.lm2m == 0, 1.0, 2.0)[keep] * np.abs(S[keep]) ** 2)
    assert abs(lhs - rhs) < 1e-12 * rhs
    S0 = S.copy()
    S0[o.lm2l == 0] = 0
    vt, vp = o.sphtor_to_spat(S0, T, lcut)
    s2, t2 = o.spat_to_sphertor(vt, vp, lcut)
    assert rel_l2(s2[keep], S0[keep]) < 1e-13 and rel_l2(t2[keep], T[keep]) < 1e-13
    assert np.all(s2[~keep] == 0) and np.all(t2[~keep] == 0)


def test_gradient_and_phi_derivative_consistency(o16):
    """grad synthesis == spheroidal part of sphtor synthesis; d/dphi checked spectrally with numpy."""
    o = o16
    rng = np.random.default_rng(3)
    S = random_spectrum(o, rng, zero_l0=True)
    gt, gp = o.scal_to_grad_spat(S, 16)
    vt, vp = o.sphtor_to_spat(S, np.zeros_like(S), 16)
    assert np.allclose(gt, vt, atol=1e-13) and np.allclose(gp, vp, atol=1e-13)
    f = o.scal_to_spat(S, 16)
    k = np.fft.rfftfreq(o.n_phi, 1.0 / o.n_phi)
    dfdphi = np.fft.irfft(1j * k[:, None] * o.minc * np.fft.rfft(f, axis=0), n=o.n_phi, axis=0)
    assert np.allclose(gp, dfdphi, atol=1e-12)
    # torpol_to_spat: radial part is the scalar synthesis of l(l+1) W
    W = random_spectrum(o, rng)
    vr, _, _ = o.torpol_to_spat(W, S, np.zeros_like(S), 16)
    assert np.allclose(vr, o.scal_to_spat(o.dLh * W, 16), atol=1e-12)


def test_toroidal_field_is_divergence_free_rotation(o16):
    """T=Y10 -> solid-body rotation: sin(theta) v_phi = -sin(theta) d/dtheta Y10 = sqrt(3/4pi) sin^2."""
    o = o16
    T = np.zeros(o.lm_max, complex)
    T[1] = 1
    vt, vp = o.sphtor_to_spat(np.zeros_like(T), T, 16)
    assert np.allclose(vt, 0, atol=1e-15)
    assert np.allclose(vp, np.sqrt(3 / 4 / np.pi) * o.sinTheta_E2[None, :], atol=1e-14)


def test_axisymmetric_transforms_match_m0_of_full_transform(o16):
    o = o16
    rng = np.random.default_rng(5)
    fl = rng.standard_normal(o.l_max + 1) + 0j
    S = np.zeros(o.lm_max, complex)
    S[: o.l_max + 1] = fl
    assert np.allclose(o.axi_to_spat(fl), o.scal_to_spat(S, 16)[0], atol=1e-13)
    ft, fp = o.toraxi_to_spat(fl, 16)
    S[0] = 0
    vt, vp = o.sphtor_to_spat(np.zeros_like(S), S, 16)
    assert np.allclose(ft, vt[0], atol=1e-13) and np.allclose(fp, vp[0], atol=1e-13)


def test_lo_map_is_a_balanced_permutation_and_transposes_invert(o16):
    o = o16
    for n_procs in (1, 2, 3, 4, 8, 9):
        lo2st, s, e = o.lo_map(n_procs)
        assert sorted(lo2st.tolist()) == list(range(o.lm_max))
        assert s[0] == 1 and e[-1] == o.lm_max and np.all(s[1:] == e[:-1] + 1)
        if n_procs <= o.l_max // 2:
            assert lo2st[0] == 0  # (l=0,m=0) first on rank 0, blocking.f90:476-484
    rng = np.random.default_rng(11)
    n_procs, n_r_max, n_fields = 4, 9, 3
    _, s, e = o.lo_map(n_procs)
    arr_LM = [rng.standard_normal((n_fields, n_r_max, e[p] - s[p] + 1)) + 0j for p in range(n_procs)]
    arr_R = o.transp_lm2r(n_procs, n_r_max, arr_LM)
    back = o.transp_r2lm(n_procs, n_r_max, arr_R)
    for p in range(n_procs):
        assert np.array_equal(back[p], arr_LM[p])
    # every (lm_st, r) holds the value of the matching lo entry
    lo2st, s, e = o.lo_map(n_procs)
    rs, re = get_blocks(n_r_max, n_procs)
    q = 2
    lm_lo = 5
    p = int(np.searchsorted(e, lm_lo + 1))
    assert arr_R[q][1, 0, lo2st[lm_lo]] == arr_LM[p][1, rs[q] - 1, lm_lo - (s[p] - 1)]


def test_phase_field_source_of_a_uniform_phase():
    """get_nl.f90:340-343 / rIter.f90:698: for a uniform phase field phi = c0 (only the (0,0) mode), no flow and no entropy
    perturbation, phiTerms is the constant -c0 (1 - c0) (phaseDiffFac (1 - 2 c0) - tmelt) / epsPhase^2 on bulk levels, so dphidt
    has that value times sqrt(4 pi) in its (0,0) entry and nothing else; boundary levels are not written by get_nl."""
    from magic_b200.workload import make_params, make_radial
    from oracle.oracle import Oracle, Params
    o, n_r = Oracle(8), 5
    p = make_params("hydro", n_r)
    op = Params()
    for n, _ in p._fields_:
        setattr(op, n, getattr(p, n))
    op.l_phase_field, op.epsPhase, op.phaseDiffFac, op.penaltyFac, op.tmelt = 1, 0.03, 1.0, 0.5, 0.11
    rad = make_radial(n_r, 8)
    f = {k: np.zeros((n_r, o.lm_max), dtype=complex) for k in ("w", "dw", "ddw", "z", "dz", "s", "phi")}
    c0 = 0.3
    f["phi"][:, 0] = c0 * np.sqrt(4 * np.pi)
    out = o.radial_loop(op, rad, f)
    want = -c0 * (1 - c0) * (1.0 * (1 - 2 * c0) - 0.11) / 0.03 ** 2 * np.sqrt(4 * np.pi)
    assert np.allclose(out["dphidt"][1:-1, 0], want, rtol=1e-13)
    assert np.abs(out["dphidt"][1:-1, 1:]).max() < 1e-12 * abs(want)
    assert not out["dphidt"][[0, -1]].any()


def test_centre_level_of_a_full_sphere_feeds_no_output():
    """v_center_sphere (nonlinear_bcs.f90:177-224) builds the l = 1 vector field at r = 0, but the centre is a boundary level
    (nBc = kbotv /= 0): get_td writes only dVxBhLM ~ r^2 there (get_td.f90: boundary branches), which vanishes at r = 0.  So no
    explicit term depends on it -- it matters to the grid outputs (graphics, movies) alone, which stay with the host.  That is
    why no golden energy series can pin it, and why nothing needs to."""
    from magic_b200.workload import make_fields, make_params, make_radial
    from oracle.oracle import Oracle, Params as OParams
    l_max, n_r = 16, 6
    o = Oracle(l_max)
    p = make_params("mhd", n_r, ktopv=1, kbotv=1)
    p.l_full_sphere = 1
    rad = {k: v.copy() for k, v in make_radial(n_r, l_max).items()}
    for k in ("r", "or1", "or2", "or4"):
        rad[k][-1] = 0.0
    f = make_fields("mhd", o.lm2l, o.lm2m, n_r, 3)
    op = OParams()
    for n, _ in p._fields_:
        setattr(op, n, getattr(p, n))
    a = o.radial_loop(op, rad, f)
    g = {k: v.copy() for k, v in f.items()}
    g["ddw"][-1, o.lm2l == 1] *= 3.7
    g["ddb"][-1, o.lm2l == 1] *= -2.1
    b = o.radial_loop(op, rad, g)
    for k, v in a.items():
        if getattr(v, "ndim", 0) == 2:
            assert not v[-1].any() and np.array_equal(v, b[k]), k
    assert np.abs(a["dwdt"][1:-1]).max() > 0
