"""End-to-end golden vectors of the reference: samples/hydro_bench_anel (BASELINE config 2).

The reference's autotest (`samples/hydro_bench_anel/unitTest.py`, rtol 1e-8) runs the anelastic hydro benchmark: polytropic
reference state over five density scale heights, stress-free walls, l_max=96, n_r_max=97 (n_cheb_max=95), l_adv_curl forced
off (u.grad u advection), viscous heating, angular-momentum correction; e_kin.TAG is logged every 10 steps.  The start state
is the conductive entropy profile plus one (l=19, m=19) mode, so the axisymmetric energy columns (1e-4 .. 1e-2 while the
total is 30 .. 300) exist ONLY through the quadratic terms of get_nl: they pin the anelastic advection branch, the
spat_to_qst analyses and get_td's anelastic scalings, which samples/dynamo_benchmark does not touch.

The Fortran host is restated in numpy (oracle/lmloop.py ShellHost: anelastic background, stress-free boundary rows,
dealiased Chebyshev solves, l_correct_AMz/AMe); the radial loop is the CPU oracle (CPU test, first logged row = 10 steps)
or the CUDA library through the C ABI (`-m gpu`, all 30 logged rows = the 300 steps of the reference run).  tests/golden/hydro_bench_anel_reference.npz holds
reference.out (tests/golden/make_hydro_bench_anel_fixture.py).
"""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
RTOL, ATOL = 1e-8, 1e-20          # samples/hydro_bench_anel/unitTest.py (magic_wizard.py:423)


@pytest.fixture(scope="module")
def golden():
    d = np.load(os.path.join(HERE, "golden", "hydro_bench_anel_reference.npz"))
    return {k: d[k] for k in d.files}


def _setup(golden, lm2l, lm2m, l_correct_AM=True):
    from magic_b200.workload import make_params, make_radial
    from oracle.lmloop import ShellHost
    n_r = int(golden["n_r_max"])
    kw = {k: float(golden[k]) for k in ("radratio", "ra", "ek", "pr", "dtmax", "alpha", "amp_s1", "strat", "polind", "g0", "g1", "g2")}
    h = ShellHost(lm2l, lm2m, None, n_r_max=n_r, n_cheb_max=int(golden["n_cheb_max"]), init_s1=int(golden["init_s1"]),
                  l_mag=False, ktopv=int(golden["ktopv"]), kbotv=int(golden["kbotv"]), l_correct_AMz=l_correct_AM, l_correct_AMe=l_correct_AM, **kw)
    p = make_params("anel", n_r, ktopv=int(golden["ktopv"]), kbotv=int(golden["kbotv"]))
    p.ViscHeatFac = h.ViscHeatFac       # DissNb * pr / ra (radial.f90:762)
    p.ra = float(golden["ra"])
    p.courfac, p.alffac = float(golden["courfac"]), float(golden["alffac"])
    l_max = int(lm2l.max())
    rad = make_radial(n_r, l_max)
    assert np.abs(rad["r"] - h.g.r).max() < 1e-15
    rad.update(rho0=h.rho0, beta=h.beta, temp0=h.temp0, orho1=1.0 / h.rho0, orho2=1.0 / h.rho0 ** 2, otemp1=1.0 / h.temp0)
    return h, p, rad


def _run(golden, h, n_rows):
    step = int(golden["n_log_step"])
    for row in range(1, n_rows + 1):
        for _ in range(step):
            h.step()
        got = np.concatenate([[h.time], h.e_kin()])
        np.testing.assert_allclose(got, golden["e_kin"][row], rtol=RTOL, atol=ATOL, err_msg=f"row {row}")


def test_reference_state_is_the_benchmark_polytrope(golden):
    """N_rho = 5: rho0(r_i) / rho0(r_o) = e^5 (radial.f90:697-712), and the start state carries no kinetic energy."""
    from oracle.oracle import Oracle, grid_sizes
    gs = grid_sizes(n_phi_tot=int(golden["n_phi_tot"]))
    assert gs["l_max"] == 96 and gs["lm_max"] == 4753
    o = Oracle(gs["l_max"], n_theta=gs["n_theta_max"], n_phi=gs["n_phi_max"], m_max=gs["m_max"])
    h, p, rad = _setup(golden, o.lm2l, o.lm2m)
    assert abs(h.rho0[-1] / h.rho0[0] / np.exp(5.0) - 1.0) < 1e-13
    assert np.all(h.e_kin() == 0.0)
    np.testing.assert_allclose(golden["e_kin"][0, 1:], 0.0)


def test_oracle_radial_loop_reproduces_reference_energies(golden):
    """CPU oracle inside the reference's time loop: the first logged row (10 steps), all 8 energy columns."""
    from oracle.oracle import Oracle, Params as OParams, grid_sizes
    gs = grid_sizes(n_phi_tot=int(golden["n_phi_tot"]))
    o = Oracle(gs["l_max"], n_theta=gs["n_theta_max"], n_phi=gs["n_phi_max"], m_max=gs["m_max"],
               threads=min(4, os.cpu_count() or 1), fast=True)  # the -O3 build of the same oracle source: ~3 s per radial loop at l_max = 96
    h, p, rad = _setup(golden, o.lm2l, o.lm2m)
    op = OParams()
    for n, _ in p._fields_:
        setattr(op, n, getattr(p, n))
    h.radial_loop = lambda f: o.radial_loop(op, rad, f)
    _run(golden, h, 1)


@pytest.mark.gpu
def test_gpu_radial_loop_reproduces_reference_energies(golden):
    """The CUDA radial loop (magic_rloop_run, host containers) inside the reference's time loop: all 30 rows (300 steps,
    e_kin grows from 31 to 322 into the nonlinear regime)."""
    from magic_b200 import RadialLoop, Sht, grid_sizes
    gs = grid_sizes(n_phi_tot=int(golden["n_phi_tot"]))
    s = Sht(gs["l_max"], m_max=gs["m_max"], n_theta_max=gs["n_theta_max"], n_phi_max=gs["n_phi_max"])
    h, p, rad = _setup(golden, s.lm2l, s.lm2m)
    rl = RadialLoop(s, p, rad)
    h.radial_loop = lambda f: rl.radialLoop(f)
    _run(golden, h, len(golden["e_kin"]) - 1)
    assert rl.launch_count() > 0
    rl.finalize()
    s.finalize_sht()
