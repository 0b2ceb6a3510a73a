"""Self-checks of the numpy restatements in oracle/rms_host.py that stand between the batches and the reference's golden series
(they are test infrastructure, but a wrong helper could mask a wrong batch): Simpson's rule on MagIC's decreasing grids, the
z-averaging on the cylindrical grid (integration.f90:157-533) on fields with known averages, the cut-back radial grid of init_rNB."""
import numpy as np

from oracle.lmloop import ChebShell
from oracle.rms_host import RmsHost, ToHost, simps


class _H:      # the few attributes ToHost / RmsHost read
    def __init__(self, n_r=33, l_max=8):
        self.g = ChebShell(n_r, 0.35, n_r - 2)
        lm = [(l, m) for m in range(l_max + 1) for l in range(m, l_max + 1)]
        self.lm2l = np.array([a for a, _ in lm])
        self.lm2m = np.array([b for _, b in lm])


def test_simps_is_exact_for_cubics_on_decreasing_grids():
    for n in (9, 10, 33, 50):
        r = np.sort(np.random.default_rng(n).random(n) * 2.0 + 0.5)[::-1]
        f = 1.0 + 2.0 * r - 0.5 * r ** 2
        exact = (r[0] - r[-1]) + (r[0] ** 2 - r[-1] ** 2) - (r[0] ** 3 - r[-1] ** 3) / 6.0
        assert abs(simps(f, r) - exact) < (1e-12 if n % 2 else 2e-2) * exact      # odd counts: exact for quadratics on any grid
    r = np.linspace(2.0, 0.5, 21)
    assert abs(simps(r ** 3, r) - (2.0 ** 4 - 0.5 ** 4) / 4.0) < 1e-12             # uniform grid: exact for cubics


def test_cylmean_of_fields_with_known_z_averages():
    h = _H()
    n_theta = 48
    theta = np.arccos(np.polynomial.legendre.leggauss(n_theta)[0])
    theta = np.sort(theta)
    T = ToHost(h, theta, toraxi_to_spat=None)
    r = h.g.r
    one = np.ones((n_theta, len(r)))
    vN, vS = T.cylmean(one)
    assert np.abs(vN - 1.0).max() < 1e-12 and np.abs(vS - 1.0).max() < 1e-12
    z = np.cos(theta)[:, None] * r[None, :]                                        # a = z: odd about the equator
    s2 = (np.sin(theta)[:, None] * r[None, :]) ** 2                                # a = s^2: constant along z
    zN, zS = T.cylmean(z)
    sN, sS = T.cylmean(s2)
    k = T.n_s_otc
    cyl = T.cyl
    assert np.abs(zN[1:k]).max() < 1e-10                                           # outside the tangent cylinder the average of z vanishes
    zmid = 0.5 * (np.sqrt(r[0] ** 2 - cyl[k:] ** 2) + np.sqrt(r[-1] ** 2 - cyl[k:] ** 2))
    assert np.abs(zN[k:] - zmid).max() < 2e-3 and np.abs(zS[k:] + zmid).max() < 2e-3   # inside: mid-height of the column, N = -S
    inner = slice(2, T.n_s_max - 2)                                                # (fourth-order interpolation; the polar-most columns extrapolate in theta)
    assert np.abs(sN[inner] - cyl[inner] ** 2).max() < 5e-4 * r[0] ** 2


def test_cut_back_grid_integrates_polynomials():
    h = _H()
    h.n_cheb_max = h.g.n_cheb_max
    R = RmsHost.__new__(RmsHost)
    h.g.n_cheb_max = h.g.n_cheb_max
    RmsHost.__init__(R, h, rCut=1e-2, rDea=0.0)
    assert (R.nCut, R.n2) == (4, 25)                                               # 33 levels, rCut = 1e-2: RMS.f90:384-411
    r = h.g.r
    rc = r[R.nCut:R.nCut + R.n2]
    # the 25 kept levels are not the Gauss-Lobatto nodes of their interval: the reference maps them through dr/dx taken
    # spectrally on the 25-point grid, which integrates smooth functions to a few 1e-5 only -- restated as it is (the golden
    # dtVrms.TAG rows are matched to 1e-9 with exactly this rule)
    for f, F in ((lambda x: x ** 2, lambda x: x ** 3 / 3.0), (lambda x: 1.0 / x ** 2, lambda x: -1.0 / x)):
        exact = F(rc[0]) - F(rc[-1])
        assert abs(R.w_cut @ f(rc) - exact) < 2e-4 * abs(exact)
    assert abs(R.w_cut.sum() - (rc[0] - rc[-1])) < 1e-12 * (rc[0] - rc[-1])      # constants: exact (sum of dr/dx weights)
    assert abs(R.volC - 4.0 / 3.0 * np.pi * (rc[0] ** 3 - rc[-1] ** 3)) < 1e-14
