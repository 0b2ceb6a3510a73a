"""End-to-end golden vectors of the reference: samples/precession.

The reference's autotest (`samples/precession/unitTest.py`, rtol 1e-8) spins up a precessing shell from rest: Po = -0.01 at
23.5 degrees, Ek = 1e-3, no buoyancy (ra = 0 switches l_heat off, Namelists.f90:429), no magnetic field, rigid walls,
l_max = 42 truncated at m_max = 5, n_r_max = n_cheb_max = 49, 200 CNAB2 steps of 1e-5, e_kin.TAG logged every 10 steps.  The
flow is driven by the Poincare force on z(1,1) in updateZ; everything else -- in particular all axisymmetric energy (columns
4-5, 8-9, 1e-13 .. 1e-2) -- exists only through the radial loop: the time-dependent precession terms PCr/PCt/PCp of get_nl
(get_nl.f90:346-357, added to the advection at rIter.f90:669-673), the Coriolis force with CorFac (1 + Po cos(alpha))
(preCalculations.f90:166) and the advection.  This pins the l_precession branch, the `time` argument of the radial loop and
a truncation with m_max < l_max.

The Fortran host is restated in numpy (oracle/lmloop.py ShellHost with l_heat off and the Poincare terms); the radial loop
is the CPU oracle (CPU test, 20 steps) or the CUDA library through the C ABI (all 200 steps).
tests/golden/precession_reference.npz holds reference.out (tests/golden/make_precession_fixture.py).
"""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
RTOL = 1e-8          # samples/precession/unitTest.py


@pytest.fixture(scope="module")
def golden():
    d = np.load(os.path.join(HERE, "golden", "precession_reference.npz"))
    return {k: d[k] for k in d.files}


def _sizes(golden):
    from magic_b200.sht import grid_sizes
    gs = grid_sizes(l_max=int(golden["l_max"]))
    assert (gs["n_theta_max"], gs["n_phi_max"]) == (64, 128)          # truncation.f90:66-73 with prime_decomposition(126)
    return gs


def _setup(golden, lm2l, lm2m):
    from magic_b200.workload import make_params, make_radial
    from oracle.lmloop import ShellHost
    n_r = int(golden["n_r_max"])
    assert len(lm2l) == 243 and int(lm2m.max()) == 5                 # sum_{m=0..5} (43 - m)
    kw = {k: float(golden[k]) for k in ("radratio", "ra", "ek", "pr", "dtmax", "alpha", "po", "prec_angle")}
    h = ShellHost(lm2l, lm2m, None, n_r_max=n_r, n_cheb_max=int(golden["n_cheb_max"]), init_s1=0, init_b1=0, l_mag=False, **kw)
    assert not h.l_heat
    p = make_params("hydro", n_r, ktopv=int(golden["ktopv"]), kbotv=int(golden["kbotv"]))
    p.l_heat = p.l_heat_nl = 0
    p.l_precession = 1                                                # po /= 0 (Namelists.f90:581-588)
    p.po, p.prec_angle = kw["po"], np.deg2rad(kw["prec_angle"])
    p.oek = 1.0 / kw["ek"]                                            # preCalculations.f90:164
    p.CorFac = p.oek * (1.0 + p.po * np.cos(p.prec_angle))            # preCalculations.f90:166
    p.ra = 0.0
    p.courfac, p.alffac = float(golden["courfac"]), float(golden["alffac"])
    rad = make_radial(n_r, int(golden["l_max"]))
    assert np.abs(rad["r"] - h.g.r).max() < 1e-15
    return h, p, rad


def _oracle_params(p):
    from oracle.oracle import Params as OParams
    op = OParams()
    for n, _ in p._fields_:
        setattr(op, n, getattr(p, n))
    return op


def _run(golden, h, n_rows):
    step = int(golden["n_log_step"])
    for row in range(1, n_rows + 1):
        for _ in range(step):
            h.step()
        got = np.concatenate([[h.time], h.e_kin()])
        np.testing.assert_allclose(got, golden["e_kin"][row], rtol=RTOL, atol=1e-30, err_msg=f"row {row}")


def _oracle(golden, gs):
    from oracle.oracle import Oracle
    return Oracle(int(golden["l_max"]), n_theta=gs["n_theta_max"], n_phi=gs["n_phi_max"], m_max=int(golden["m_max"]),
                  threads=min(4, os.cpu_count() or 1))


def test_oracle_radial_loop_reproduces_reference_energies(golden):
    """CPU oracle inside the reference's time loop: the first two logged rows (20 steps); the start is at rest."""
    gs = _sizes(golden)
    o = _oracle(golden, gs)
    h, p, rad = _setup(golden, o.lm2l, o.lm2m)
    assert np.all(h.e_kin() == 0.0) and np.all(golden["e_kin"][0] == 0.0)
    op = _oracle_params(p)
    h.radial_loop = lambda f: o.radial_loop(op, rad, f, time=h.time)   # timeStage of radialLoopG = time of the fields
    _run(golden, h, 2)


def test_the_axisymmetric_energies_come_from_the_precession_terms(golden):
    """Negative control: without PCr/PCt/PCp in get_nl the axisymmetric columns are wrong by O(1) after ten steps, and
    with a radial loop that ignores `time` they are off too (measured at row 5: 0.88 / 0.97 relative)."""
    gs = _sizes(golden)
    o = _oracle(golden, gs)
    for kind in ("no_pc", "no_time"):
        h, p, rad = _setup(golden, o.lm2l, o.lm2m)
        op = _oracle_params(p)
        if kind == "no_pc":
            op.l_precession = 0
        h.radial_loop = (lambda f: o.radial_loop(op, rad, f, time=h.time)) if kind == "no_pc" else \
            (lambda f: o.radial_loop(op, rad, f, time=0.5e-3))
        for _ in range(int(golden["n_log_step"])):
            h.step()
        dev = np.abs(h.e_kin() / golden["e_kin"][1][1:] - 1.0)
        assert dev[2] > 1e-3, (kind, dev)          # e_kin_pol_axi


@pytest.mark.gpu
def test_gpu_radial_loop_reproduces_reference_energies(golden):
    """The CUDA radial loop (magic_rloop_run, host containers) inside the reference's time loop: all 20 logged rows."""
    from magic_b200 import RadialLoop, Sht
    gs = _sizes(golden)
    s = Sht(int(golden["l_max"]), m_max=int(golden["m_max"]), n_theta_max=gs["n_theta_max"], n_phi_max=gs["n_phi_max"])
    h, p, rad = _setup(golden, s.lm2l, s.lm2m)
    rl = RadialLoop(s, p, rad)
    h.radial_loop = lambda f: rl.radialLoop(f, time=h.time)
    _run(golden, h, len(golden["e_kin"]) - 1)
    assert rl.launch_count() > 0
    rl.finalize()
    s.finalize_sht()
