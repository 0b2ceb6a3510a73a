"""GPU parity: the 17 `module sht` procedures through the C ABI (host buffers) vs the CPU oracle.

Bar (BASELINE.json north_star): relative L2 <= 1e-12 per transform; degrees above lcut exact zeros.
"""
import numpy as np
import pytest

from tests.util import random_spectrum, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-12

CASES = {
    "l16": dict(l_max=16),
    "l32_minc3": dict(l_max=0, n_phi_tot=96, minc=3),
    "l96": dict(l_max=0, n_phi_tot=288),
    "l21_odd": dict(l_max=21),
}


@pytest.fixture(scope="module", params=list(CASES))
def ctx(request):
    from magic_b200 import Sht, grid_sizes
    from oracle.oracle import Oracle
    c = CASES[request.param]
    gs = grid_sizes(l_max=c["l_max"], n_phi_tot=c.get("n_phi_tot", 0), minc=c.get("minc", 1))
    o = Oracle(gs["l_max"], minc=c.get("minc", 1), n_theta=gs["n_theta_max"], n_phi=gs["n_phi_max"], m_max=gs["m_max"])
    s = Sht(gs["l_max"], m_max=gs["m_max"], minc=c.get("minc", 1), n_theta_max=gs["n_theta_max"], n_phi_max=gs["n_phi_max"])
    yield o, s
    s.finalize_sht()


def lcuts(o):
    return [o.l_max, max(1, (2 * o.l_max) // 3)]


def test_grid_matches_oracle(ctx):
    o, s = ctx
    th, g = s.get_grid()
    assert s.l_scrambled_theta
    assert np.allclose(th, o.theta_ord, rtol=0, atol=1e-15)
    assert np.allclose(g[: o.n_theta // 2], o.gauss[0::2], rtol=1e-15)


def test_scal_to_spat(ctx):
    o, s = ctx
    rng = np.random.default_rng(1)
    for lcut in lcuts(o):
        S = random_spectrum(o, rng)
        assert rel_l2(s.scal_to_spat(S, lcut), o.scal_to_spat(S, lcut)) < TOL


def test_grad_variants(ctx):
    o, s = ctx
    rng = np.random.default_rng(2)
    for lcut in lcuts(o):
        S = random_spectrum(o, rng)
        for name in ["scal_to_grad_spat", "pol_to_grad_spat"]:
            a, b = getattr(s, name)(S, lcut)
            ra, rb = getattr(o, name)(S, lcut)
            assert rel_l2(a, ra) < TOL and rel_l2(b, rb) < TOL, name


def test_torpol_to_spat_and_sphtor(ctx):
    o, s = ctx
    rng = np.random.default_rng(3)
    for lcut in lcuts(o):
        W, dW, Z = (random_spectrum(o, rng, zero_l0=True) for _ in range(3))
        got = s.torpol_to_spat(W, dW, Z, lcut)
        ref = o.torpol_to_spat(W, dW, Z, lcut)
        for a, b in zip(got, ref):
            assert rel_l2(a, b) < TOL
        got = s.sphtor_to_spat(dW, Z, lcut)
        ref = o.sphtor_to_spat(dW, Z, lcut)
        for a, b in zip(got, ref):
            assert rel_l2(a, b) < TOL
        got = s.torpol_to_dphspat(dW, Z, lcut)
        ref = o.torpol_to_dphspat(dW, Z, lcut)
        for a, b in zip(got, ref):
            assert rel_l2(a, b) < TOL


def test_curl_variants(ctx):
    o, s = ctx
    rng = np.random.default_rng(4)
    for lcut in lcuts(o):
        B, ddB, J, dJ = (random_spectrum(o, rng, zero_l0=True) for _ in range(4))
        got = s.torpol_to_curl_spat(1.37, B, ddB, J, dJ, lcut)
        ref = o.torpol_to_curl_spat(1.37, B, ddB, J, dJ, lcut)
        for a, b in zip(got, ref):
            assert rel_l2(a, b) < TOL
        assert rel_l2(s.pol_to_curlr_spat(J, lcut), o.pol_to_curlr_spat(J, lcut)) < TOL


def test_inner_core_variants(ctx):
    o, s = ctx
    rng = np.random.default_rng(5)
    W, dW, Z, dJ = (random_spectrum(o, rng, zero_l0=True) for _ in range(4))
    for a, b in zip(s.torpol_to_spat_IC(0.3, 0.5385, W, dW, Z), o.torpol_to_spat_IC(0.3, 0.5385, W, dW, Z)):
        assert rel_l2(a, b) < TOL
    for a, b in zip(s.torpol_to_curl_spat_IC(0.3, 0.5385, W, dW, Z, dJ), o.torpol_to_curl_spat_IC(0.3, 0.5385, W, dW, Z, dJ)):
        assert rel_l2(a, b) < TOL


def test_analysis(ctx):
    o, s = ctx
    rng = np.random.default_rng(6)
    f, g, h = (rng.standard_normal((o.n_phi, o.n_theta)) for _ in range(3))
    for lcut in lcuts(o):
        ref = o.scal_to_SH(f, lcut)
        got = s.scal_to_SH(f, lcut)
        assert rel_l2(got, ref) < TOL
        assert np.all(got[o.lm2l > lcut] == 0)
        for a, b in zip(s.spat_to_qst(f, g, h, lcut), o.spat_to_qst(f, g, h, lcut)):
            assert rel_l2(a, b) < TOL
            assert np.all(a[o.lm2l > lcut] == 0)
        for a, b in zip(s.spat_to_sphertor(g, h, lcut), o.spat_to_sphertor(g, h, lcut)):
            assert rel_l2(a, b) < TOL


def test_per_coefficient_roundtrip(ctx):
    """analysis(synthesis(S)) == S per coefficient: |diff| <= 1e-12 * max|S| (size-independent property)."""
    o, s = ctx
    rng = np.random.default_rng(7)
    S = random_spectrum(o, rng)
    T = random_spectrum(o, rng, zero_l0=True)
    S2 = s.scal_to_SH(s.scal_to_spat(S, o.l_max), o.l_max)
    assert np.max(np.abs(S2 - S)) < 1e-12 * np.max(np.abs(S))
    S0 = S.copy()
    S0[o.lm2l == 0] = 0
    vt, vp = s.sphtor_to_spat(S0, T, o.l_max)
    s2, t2 = s.spat_to_sphertor(vt, vp, o.l_max)
    assert np.max(np.abs(s2 - S0)) < 1e-12 * np.max(np.abs(S0))
    assert np.max(np.abs(t2 - T)) < 1e-12 * np.max(np.abs(T))


def test_axisymmetric(ctx):
    o, s = ctx
    rng = np.random.default_rng(8)
    fl = rng.standard_normal(o.l_max + 1) + 1j * rng.standard_normal(o.l_max + 1)
    assert rel_l2(s.axi_to_spat(fl), o.axi_to_spat(fl)) < TOL
    for lcut in lcuts(o):
        for a, b in zip(s.toraxi_to_spat(fl, lcut), o.toraxi_to_spat(fl, lcut)):
            assert np.linalg.norm(a - b) < TOL * max(np.linalg.norm(b), 1.0)


def test_bitwise_stable_across_runs(ctx):
    o, s = ctx
    rng = np.random.default_rng(9)
    W, dW, Z = (random_spectrum(o, rng, zero_l0=True) for _ in range(3))
    a = s.torpol_to_spat(W, dW, Z, o.l_max)
    b = s.torpol_to_spat(W, dW, Z, o.l_max)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    q1 = s.spat_to_qst(a[0], a[1], a[2], o.l_max)
    q2 = s.spat_to_qst(a[0], a[1], a[2], o.l_max)
    for x, y in zip(q1, q2):
        assert np.array_equal(x, y)


def test_nlat_padded_and_errors():
    from magic_b200 import MagicError, Sht
    from oracle.oracle import Oracle
    o = Oracle(16)
    s = Sht(16, nlat_padded=32)
    rng = np.random.default_rng(10)
    S = random_spectrum(o, rng)
    f = s.scal_to_spat(S, 16)
    assert f.shape == (48, 32)
    assert rel_l2(f[:, :24], o.scal_to_spat(S, 16)) < TOL
    assert np.all(f[:, 24:] == 0)  # padded rows carry zeros (horizontal.f90:72-77)
    fin = np.zeros((48, 32))
    fin[:, :24] = o.scal_to_spat(S, 16)
    assert rel_l2(s.scal_to_SH(fin, 16), S) < 1e-12
    with pytest.raises(MagicError):
        s.scal_to_spat(S, 17)
    with pytest.raises(MagicError):
        Sht(16, n_theta_max=22, n_phi_max=48)
    s.finalize_sht()
