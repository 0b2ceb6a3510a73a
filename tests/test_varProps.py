"""End-to-end golden vectors of the reference: samples/varProps (Chebyshev stage).

The reference's autotest (`samples/varProps/unitTest.py`, rtol 1e-8) runs an anelastic hydro case over three density scale
heights whose kinematic viscosity and thermal diffusivity vary with radius as rho^-1/2 (nVarVisc = nVarDiff = 2, difExp =
-0.5): polytropic index 2, gravity ~ 1/r^2, stress-free walls, l_max = 32, n_r_max = n_cheb_max = 33, 250 CNAB2 steps of
1e-4 from init_s1 = 707, e_kin.TAG logged every 10 steps (the first 26 rows of reference.out; the rest repeats the case on
a mapped and on a finite-difference grid).  On the radial-loop side the variable viscosity enters the viscous heating of
get_nl (get_nl.f90:402-425, the `visc` entry of magic_radial): with visc = 1 there the axisymmetric energies are off by
5e-2 after 100 steps.  Everything else the profiles touch is in the LM-side host (oracle/lmloop.py: dLvisc, ddLvisc,
kappa, dLkappa in the matrices and implicit terms of updateS / updateZ / updateWP and in the conductive start state).

The radial loop is the CPU oracle (CPU test, 20 steps) or the CUDA library through the C ABI (250 steps).
tests/golden/varProps_reference.npz holds the 26 rows (tests/golden/make_varProps_fixture.py).
"""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
RTOL, ATOL = 1e-8, 1e-20          # samples/varProps/unitTest.py


@pytest.fixture(scope="module")
def golden():
    d = np.load(os.path.join(HERE, "golden", "varProps_reference.npz"))
    return {k: d[k] for k in d.files}


def _sizes(golden):
    from magic_b200.sht import grid_sizes
    gs = grid_sizes(n_phi_tot=int(golden["n_phi_tot"]))
    assert (gs["l_max"], gs["lm_max"]) == (32, 561)
    return gs


def _setup(golden, lm2l, lm2m):
    from magic_b200.workload import make_params, make_radial
    from oracle.lmloop import ShellHost
    n_r = int(golden["n_r_max"])
    kw = {k: float(golden[k]) for k in ("radratio", "ra", "ek", "pr", "dtmax", "alpha", "amp_s1", "strat", "polind", "g0", "g1", "g2")}
    h = ShellHost(lm2l, lm2m, None, n_r_max=n_r, n_cheb_max=int(golden["n_cheb_max"]), init_s1=int(golden["init_s1"]), l_mag=False,
                  ktopv=int(golden["ktopv"]), kbotv=int(golden["kbotv"]), dif_exp=float(golden["difExp"]), **kw)
    assert abs(h.rho0[-1] / h.rho0[0] / np.exp(3.0) - 1.0) < 1e-13
    assert h.visc[-1] == 1.0 and abs(h.visc[0] - np.exp(1.5)) < 1e-12      # (rho0 / rho0(icb))^-1/2 (radial.f90 nVarVisc = 2)
    p = make_params("anel", n_r, ktopv=int(golden["ktopv"]), kbotv=int(golden["kbotv"]))
    p.ViscHeatFac, p.ra = h.ViscHeatFac, kw["ra"]
    p.courfac, p.alffac = float(golden["courfac"]), float(golden["alffac"])
    rad = make_radial(n_r, 32)
    assert np.abs(rad["r"] - h.g.r).max() < 1e-15
    rad.update(rho0=h.rho0, beta=h.beta, temp0=h.temp0, orho1=1.0 / h.rho0, orho2=1.0 / h.rho0 ** 2, otemp1=1.0 / h.temp0,
               visc=h.visc)
    return h, p, rad


def _run(golden, h, n_rows):
    for row in range(1, n_rows + 1):
        for _ in range(int(golden["n_log_step"])):
            h.step()
        got = np.concatenate([[h.time], h.e_kin()])
        np.testing.assert_allclose(got, golden["e_kin"][row], rtol=RTOL, atol=ATOL, err_msg=f"row {row}")


def _oracle_host(golden, tweak=None):
    from oracle.oracle import Oracle, Params as OParams
    gs = _sizes(golden)
    o = Oracle(gs["l_max"], n_theta=gs["n_theta_max"], n_phi=gs["n_phi_max"], m_max=gs["m_max"], threads=min(4, os.cpu_count() or 1))
    h, p, rad = _setup(golden, o.lm2l, o.lm2m)
    op = OParams()
    for n, _ in p._fields_:
        setattr(op, n, getattr(p, n))
    if tweak:
        tweak(rad)
    h.radial_loop = lambda f: o.radial_loop(op, rad, f)
    return h


def test_oracle_radial_loop_reproduces_reference_energies(golden):
    """CPU oracle inside the reference's time loop: the first two logged rows (20 steps) from the conductive start state."""
    h = _oracle_host(golden)
    assert np.all(h.e_kin() == 0.0) and np.all(golden["e_kin"][0] == 0.0)
    _run(golden, h, 2)


def test_the_energies_see_the_viscosity_profile_in_the_loop(golden):
    """Negative control: visc = 1 in the loop's viscous heating (host unchanged) moves the axisymmetric columns, which here
    are tiny and purely nonlinear, by more than 1e-6 within ten steps."""
    h = _oracle_host(golden, tweak=lambda rad: rad.__setitem__("visc", np.ones_like(rad["visc"])))
    for _ in range(int(golden["n_log_step"])):
        h.step()
    dev = np.abs(h.e_kin() / golden["e_kin"][1][1:] - 1.0)
    assert max(dev[2], dev[3]) > 1e-6, dev


@pytest.mark.gpu
def test_gpu_radial_loop_reproduces_reference_energies(golden):
    """The CUDA radial loop (magic_rloop_run, host containers) inside the reference's time loop: all 25 logged rows."""
    from magic_b200 import RadialLoop, Sht
    gs = _sizes(golden)
    s = Sht(gs["l_max"], m_max=gs["m_max"], n_theta_max=gs["n_theta_max"], n_phi_max=gs["n_phi_max"])
    h, p, rad = _setup(golden, s.lm2l, s.lm2m)
    rl = RadialLoop(s, p, rad)
    h.radial_loop = lambda f: rl.radialLoop(f)
    _run(golden, h, len(golden["e_kin"]) - 1)
    assert rl.launch_count() > 0
    rl.finalize()
    s.finalize_sht()
