"""Small independent checks of the building blocks of the host restatements (oracle/lmloop.py, oracle/lmloop_fd.py).  The
end-to-end golden runs validate them together; these pin each one on its own against a closed form."""
import numpy as np


def test_fornberg_weights_differentiate_polynomials_exactly():
    from oracle.lmloop_fd import fd_weights
    rng = np.random.default_rng(3)
    x = np.sort(rng.uniform(-0.3, 0.4, 7))
    c = fd_weights(x, 4)                      # weights at 0 on 7 arbitrary nodes, derivatives 0..4
    for deg in range(7):
        f = x ** deg
        for k in range(5):
            exact = float(np.prod(np.arange(deg, deg - k, -1))) if deg == k else 0.0   # d^k/dx^k x^deg at 0
            assert abs(c[:, k] @ f - exact) < 1e-7 * max(1.0, np.abs(c[:, k]).max()), (deg, k)


def test_chebyshev_shell_truncated_derivative():
    from oracle.lmloop import ChebShell
    g = ChebShell(33, 0.35)
    r = g.r
    assert np.abs(g.D1 @ r ** 3 - 3 * r ** 2).max() < 1e-10 and np.abs(g.D2 @ r ** 3 - 6 * r).max() < 1e-8
    assert np.abs(g.D1t - g.D1).max() == 0.0                       # n_cheb_max = n_r_max: nothing is cut
    assert abs(g.rInt_R(r ** 2) - (g.r_cmb ** 3 - g.r_icb ** 3) / 3.0) < 1e-13
    h = ChebShell(33, 0.35, n_cheb_max=31)
    top = h.T[:, 32]                                               # the highest mode alone
    assert np.abs(h.D1t @ top).max() < 1e-10 and np.abs(h.D1 @ top).max() > 1.0
    assert np.abs(h.D1t @ r ** 3 - 3 * r ** 2).max() < 1e-10       # low modes are untouched


def test_even_chebyshev_inner_core_basis():
    from oracle.lmloop import ChebEvenIC
    ri = 7.0 / 13.0
    ic = ChebEvenIC(17, 15, ri)
    r = ic.r
    assert r[0] == ri and r[-1] == 0.0 and np.all(np.diff(r) < 0)
    f = 1.0 + 0.5 * r ** 2 - 2.0 * r ** 4
    assert np.abs(ic.D1 @ f - (r - 8.0 * r ** 3)).max() < 1e-12
    assert np.abs(ic.D2 @ f - (1.0 - 24.0 * r ** 2)).max() < 1e-10
    c = np.linalg.solve(ic.B0, f)                                  # even modes T_0, T_2, T_4 only
    assert np.abs(c[3:]).max() < 1e-13
    assert abs((ic.D1 @ f)[-1]) < 1e-12                            # an even function has no slope at the centre


def test_host_state_round_trip_and_dirk_tables():
    from oracle.lmloop import DirkShellHost, ShellHost
    from oracle.oracle import Oracle
    o = Oracle(8)
    h = ShellHost(o.lm2l, o.lm2m, None, n_r_max=17, l_mag=True, l_cond_ic=True, l_rot_ic=True, n_r_ic_max=9, n_cheb_ic_max=9)
    rng = np.random.default_rng(0)
    h.w += 1e-3 * rng.standard_normal(h.w.shape)
    h.omega_ic, h.time = 1.5, 0.25
    h.expl["z"][1] = rng.standard_normal(h.z.shape) + 0j
    d = h.state_dict()
    g = ShellHost(o.lm2l, o.lm2m, None, n_r_max=17, l_mag=True, l_cond_ic=True, l_rot_ic=True, n_r_ic_max=9, n_cheb_ic_max=9)
    g.load_state_dict(d)
    assert np.array_equal(g.w, h.w) and np.array_equal(g.b_ic, h.b_ic) and np.array_equal(g.expl["z"][1], h.expl["z"][1])
    assert (g.omega_ic, g.time) == (1.5, 0.25)
    # BPR353 (dirk_schemes.f90:723-742): consistent rows (sum of the implicit row = sum of the explicit row = c), SDIRK diagonal
    k = DirkShellHost(o.lm2l, o.lm2m, None, n_r_max=17, l_mag=False)
    for i in range(1, 5):
        assert abs(k.a_imp[i].sum() - k.a_exp[i].sum()) < 1e-15 and abs(k.a_imp[i].sum() - k.c_stage[i - 1]) < 1e-15
        assert k.a_imp[i, i] == 0.5
    assert np.array_equal(k.a_exp[4], k.a_exp[3]) and not k.l_exp_calc[3]      # stage 4 needs no new explicit term
