"""End-to-end golden vectors of the reference: samples/varCond.

The reference's autotest (`samples/varCond/unitTest.py`, rtol 1e-8) runs an ANELASTIC dynamo: one density scale height
(polytropic index 2, gravity ~ r), radratio 0.2, Ra = 1.5e7, Ek = 1e-4, Pm = 1, an electrical conductivity that drops by
a factor 40 towards the surface (nVarCond = 2), a conducting non-rotating inner core, stress-free outer and rigid inner
wall, l_correct_AMz / AMe; l_max = 32, n_r_max = 49 (n_cheb_max = 47), CNAB2 with dt = 5e-6 from init_s1 = 505 / init_b1 =
3, e_kin.TAG and e_mag_oc.TAG logged every 10 steps.  The field is strong (magnetic energy 4e5 against a kinetic energy of
7e2 after 10 steps), so the run is driven by the Lorentz force from step one.  On the radial-loop side this is the only
pinned case with a magnetic field in an anelastic background: u.grad u advection together with the Lorentz force scaled
by 1/rho, Ohmic heating with lambda(r) in the entropy equation (get_nl.f90:426-432), viscous heating, the induction
term with orho1, lMagNlBc through the stress-free top AND the conducting inner core (boundary levels with nBc = 1 and 2).

Host: oracle/lmloop.py ShellHost (anelastic background, conducting inner core, variable conductivity in get_bMat /
get_mag_rhs_imp).  The radial loop is the CPU oracle (CPU test, 20 steps) or the CUDA library through the C ABI (500 steps).
tests/golden/varCond_reference.npz holds reference.out / referenceMag.out (tests/golden/make_varCond_fixture.py).
"""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
RTOL, ATOL = 1e-8, 1e-20          # samples/varCond/unitTest.py


@pytest.fixture(scope="module")
def golden():
    d = np.load(os.path.join(HERE, "golden", "varCond_reference.npz"))
    return {k: d[k] for k in d.files}


def _sizes(golden):
    from magic_b200.sht import grid_sizes
    gs = grid_sizes(n_phi_tot=int(golden["n_phi_tot"]))
    assert (gs["l_max"], gs["lm_max"]) == (32, 561)
    return gs


def _setup(golden, lm2l, lm2m):
    from magic_b200.workload import make_params
    from oracle.lmloop import ShellHost
    N = int(golden["n_r_max"])
    kw = {k: float(golden[k]) for k in ("radratio", "ra", "ek", "pr", "prmag", "dtmax", "alpha", "amp_s1", "amp_b1", "strat",
                                        "polind", "g0", "g1", "g2", "sigma_ratio")}
    h = ShellHost(lm2l, lm2m, None, n_r_max=N, n_cheb_max=int(golden["n_cheb_max"]), init_s1=int(golden["init_s1"]),
                  init_b1=int(golden["init_b1"]), l_mag=True, ktopv=int(golden["ktopv"]), kbotv=int(golden["kbotv"]),
                  l_correct_AMz=True, l_correct_AMe=True, l_cond_ic=True, l_rot_ic=False,
                  n_r_ic_max=int(golden["n_r_ic_max"]), n_cheb_ic_max=int(golden["n_cheb_ic_max"]),
                  var_cond={k: float(golden[k]) for k in ("con_DecRate", "con_RadRatio", "con_LambdaMatch")}, **kw)
    assert abs(h.rho0[-1] / h.rho0[0] / np.e - 1.0) < 1e-13            # N_rho = 1
    assert h.lam[-1] == 1.0 and 40.0 < h.lam[0] < 40.5                 # lambda = 1 / sigma (radial.f90:903-916)
    p = make_params("anel", N, ktopv=int(golden["ktopv"]), kbotv=int(golden["kbotv"]))
    p.l_mag = p.l_mag_nl = p.l_mag_LF = 1
    p.l_cond_ic = 1
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
    lR = np.full(N, 32)
    rad = dict(nR=np.arange(1, N + 1, dtype=np.int32), l_R=lR.astype(np.int32), r=r, or1=g.or1, or2=g.or2, or4=g.or2 ** 2,
               orho1=1.0 / h.rho0, orho2=1.0 / h.rho0 ** 2, beta=h.beta, rho0=h.rho0, otemp1=1.0 / h.temp0, temp0=h.temp0,
               visc=one, epscProf=one, delxr2=delxr2, delxh2=r ** 2 / (lR * (lR + 1.0)))
    rad["lambda"] = h.lam
    return h, p, rad


def _check(golden, h, row):
    gk = np.concatenate([[h.time], h.e_kin()])
    gm = np.concatenate([[h.time], h.e_mag_oc()])
    np.testing.assert_allclose(gk, golden["e_kin"][row], rtol=RTOL, atol=ATOL, err_msg=f"e_kin row {row}")
    np.testing.assert_allclose(gm, golden["e_mag_oc"][row], rtol=RTOL, atol=ATOL, err_msg=f"e_mag_oc row {row}")


def _run(golden, h, n_rows):
    for row in range(1, n_rows + 1):
        for _ in range(int(golden["n_log_step"])):
            h.step()
        _check(golden, h, row)


def _oracle_host(golden, tweak=None):
    from oracle.oracle import Oracle, Params as OParams
    gs = _sizes(golden)
    o = Oracle(gs["l_max"], n_theta=gs["n_theta_max"], n_phi=gs["n_phi_max"], m_max=gs["m_max"], threads=min(4, os.cpu_count() or 1))
    h, p, rad = _setup(golden, o.lm2l, o.lm2m)
    op = OParams()
    for n, _ in p._fields_:
        setattr(op, n, getattr(p, n))
    if tweak:
        tweak(op, rad)
    h.radial_loop = lambda f: o.radial_loop(op, rad, f)
    return h


def test_oracle_radial_loop_reproduces_reference_energies(golden):
    """CPU oracle inside the reference's time loop: row 0 (start fields with a conducting inner core) and the first two
    logged rows (20 steps), 8 kinetic and 12 magnetic energy columns."""
    h = _oracle_host(golden)
    _check(golden, h, 0)
    _run(golden, h, 2)


def test_the_energies_see_the_anelastic_magnetic_terms(golden):
    """Negative controls after ten steps: a uniform conductivity in the loop's Ohmic heating (lambda = 1) and a loop without
    Ohmic heating at all (OhmLossFac = 0) both leave the reference."""
    for tweak, col, floor in ((lambda op, rad: rad.__setitem__("lambda", np.ones_like(rad["lambda"])), 1, 1e-7),
                              (lambda op, rad: setattr(op, "OhmLossFac", 0.0), 1, 1e-7)):
        h = _oracle_host(golden, tweak)
        for _ in range(int(golden["n_log_step"])):
            h.step()
        dev = np.abs(h.e_kin() / golden["e_kin"][1][1:] - 1.0)
        assert dev.max() > floor, dev


@pytest.mark.gpu
def test_gpu_radial_loop_reproduces_reference_energies(golden):
    """The CUDA radial loop (magic_rloop_run, host containers) inside the reference's time loop: all 50 logged rows."""
    from magic_b200 import RadialLoop, Sht
    gs = _sizes(golden)
    s = Sht(gs["l_max"], m_max=gs["m_max"], n_theta_max=gs["n_theta_max"], n_phi_max=gs["n_phi_max"])
    h, p, rad = _setup(golden, s.lm2l, s.lm2m)
    rl = RadialLoop(s, p, rad)
    h.radial_loop = lambda f: rl.radialLoop(f)
    _check(golden, h, 0)
    _run(golden, h, len(golden["e_kin"]) - 1)
    assert rl.launch_count() > 0
    rl.finalize()
    s.finalize_sht()
