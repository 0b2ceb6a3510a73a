"""Golden vectors of the reference: samples/testTruncations (26 grids, one CNAB2 step each).

`samples/testTruncations/unitTest.py` runs a nearly Boussinesq anelastic hydro case (strat=0.01, stress-free walls,
n_r_max=65, n_cheb_max=63) on 26 horizontal truncations -- n_phi_tot in {96 ... 1024} with minc 1 and 4, i.e. FFT lengths
with factors 2, 3 and 5, even and odd l_max, azimuthal symmetry -- and compares e_kin.TAG after the first step with
reference.out at rtol 1e-8.  The flow starts from rest, so the expected energy (1.25636584e-3, poloidal only) is the same
on every grid and the radial loop must contribute exact-to-rounding zeros on each of them: a garbage coefficient from any
transform size, symmetry or lcut path would show up in the 9 printed digits.  The CUDA library runs all 26 grids
(`-m gpu`); the CPU oracle the four smallest.
"""
import numpy as np
import pytest

from tests.test_hydro_bench_anel import _setup

N_PHI = [96, 96, 128, 128, 192, 192, 256, 256, 288, 288, 320, 320, 384, 384, 400, 400, 512, 512, 640, 640, 768, 768, 800, 800,
         864, 1024]                                   # samples/testTruncations/unitTest.py:72-76
MINC = [1, 4, 1, 4, 1, 4, 1, 4, 1, 4, 1, 4, 1, 4, 1, 4, 1, 4, 1, 4, 1, 4, 1, 4, 4, 4]
E_KIN_ROW1 = np.array([1e-4, 1.25636584e-3, 0.0, 0.0, 0.0, 1.25636584e-3, 0.0, 0.0, 0.0])   # every odd row of reference.out
INPUT = dict(n_r_max=65, n_cheb_max=63, ra=1.1e5, ek=1e-3, pr=1.0, strat=1e-2, polind=2.0, radratio=0.35, g0=0.0, g1=1.0, g2=0.0,
             dtmax=1e-4, alpha=0.6, init_s1=404, amp_s1=0.01, ktopv=1, kbotv=1, courfac=2.5, alffac=1.0)   # input.nml


def _check(h):
    h.step()
    got = np.concatenate([[h.time], h.e_kin()])
    # columns that are exactly zero in the reference: the printed 0.00000000E+00 allows anything below 5e-9 of the format's
    # unit; demand that the nonlinear terms left them 12 orders of magnitude below the poloidal energy
    np.testing.assert_allclose(got[[1, 5]], E_KIN_ROW1[[1, 5]], rtol=1e-8)
    assert np.all(np.abs(got[[2, 3, 4, 6, 7, 8]]) < 1e-15 * E_KIN_ROW1[1]), got


def _host(lm2l, lm2m):
    h, p, rad = _setup(INPUT, lm2l, lm2m, l_correct_AM=False)
    return h, p, rad


@pytest.mark.parametrize("k", range(4))
def test_oracle_one_step(k):
    from oracle.oracle import Oracle, Params as OParams, grid_sizes
    gs = grid_sizes(n_phi_tot=N_PHI[k], minc=MINC[k])
    o = Oracle(gs["l_max"], minc=MINC[k], n_theta=gs["n_theta_max"], n_phi=gs["n_phi_max"], m_max=gs["m_max"], threads=4)
    h, p, rad = _host(o.lm2l, o.lm2m)
    op = OParams()
    for n, _ in p._fields_:
        setattr(op, n, getattr(p, n))
    h.radial_loop = lambda f: o.radial_loop(op, rad, f)
    _check(h)


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(len(N_PHI)))
def test_gpu_one_step(k):
    from magic_b200 import RadialLoop, Sht, grid_sizes
    gs = grid_sizes(n_phi_tot=N_PHI[k], minc=MINC[k])
    s = Sht(gs["l_max"], m_max=gs["m_max"], minc=MINC[k], n_theta_max=gs["n_theta_max"], n_phi_max=gs["n_phi_max"])
    h, p, rad = _host(s.lm2l, s.lm2m)
    rl = RadialLoop(s, p, rad)
    h.radial_loop = lambda f: rl.radialLoop(f)
    _check(h)
    rl.finalize()
    s.finalize_sht()
