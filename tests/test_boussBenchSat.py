"""End-to-end golden vectors of the reference: samples/boussBenchSat, time-stepped.

tests/test_reference_energy.py uses this sample's checkpoint as a static golden vector (kinetic energy of the stored state).
Here the reference's autotest itself (`samples/boussBenchSat/unitTest.py`, rtol 1e-8) is replayed: the SATURATED benchmark
dynamo (Christensen et al. case 2: conducting and freely rotating inner core, sigma_ratio = 1, rigid walls; l_max = 64 with
minc = 4, n_r_max = 33 / n_cheb_max = 31, inner core 17 / 15) is restarted from `checkpoint_end.start` and advanced by 25 steps
of the IMEX Runge-Kutta scheme BPR353 (dt = 2e-4, three radial loops per step); e_kin.TAG is logged every 5 steps.  The
dynamo drifts steadily, so the six rows are equal to the nine printed digits and the inner core keeps omega_ic = -2.6578397;
a loop whose Lorentz force, induction, moving-wall terms or torque were off would leave that state at once.  This is the
pinned case where everything acts together on a real, fully nonlinear MHD state: advection, Lorentz force, induction,
entropy advection, Coriolis couplings, lMagNlBc boundary levels, v_rigid_boundary with omega_ic, the Lorentz torque.

Host: oracle/lmloop.py DirkShellHost with l_cond_ic / l_rot_ic.  The radial loop is the CPU oracle (CPU test, 5 steps) or the
CUDA library through the C ABI (25 steps).  tests/golden/boussBenchSat_ckpt.npz holds the checkpoint fields (incl. pressure and
inner-core potentials), omega_ic and reference.out (tests/golden/make_checkpoint_fixture.py).
"""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
RTOL = 1e-8          # samples/boussBenchSat/unitTest.py


@pytest.fixture(scope="module")
def golden():
    d = np.load(os.path.join(HERE, "golden", "boussBenchSat_ckpt.npz"))
    return {k: d[k] for k in d.files}


def _sizes(golden):
    from magic_b200.sht import grid_sizes
    gs = grid_sizes(n_phi_tot=int(golden["n_phi_tot"]), minc=int(golden["minc"]))
    assert (gs["l_max"], gs["lm_max"]) == (64, 561)
    return gs


class _Ckpt:
    def __init__(self, golden, lm_max):
        self.r, self.lm_max, self.time = golden["radius"], lm_max, float(golden["time"])
        self.fields = {k: golden[k] for k in ("w", "z", "p", "s", "b", "aj", "b_ic", "aj_ic")}
        self.past, self.scalars_past, self.rotation = {}, {}, {"omega_ic1": float(golden["omega_ic1"])}


def _setup(golden, lm2l, lm2m):
    from magic_b200.workload import make_params, make_radial
    from oracle.lmloop import DirkShellHost
    n_r = int(golden["n_r_max"])
    h = DirkShellHost(lm2l, lm2m, None, n_r_max=n_r, n_cheb_max=31, radratio=float(golden["radratio"]), ra=1.1e5,
                      ek=float(golden["ek"]), pr=1.0, prmag=float(golden["prmag"]), dtmax=float(golden["dt"][0]), alpha=0.6,
                      init_s1=0, init_b1=0, l_mag=True, l_cond_ic=True, l_rot_ic=True, sigma_ratio=1.0, n_r_ic_max=17,
                      n_cheb_ic_max=15, time_scheme="BPR353")
    h.load_checkpoint(_Ckpt(golden, len(lm2l)))
    p = make_params("mhd", n_r)
    p.l_cond_ic = p.l_rot_ic = 1
    p.ra = 1.1e5
    p.courfac, p.alffac = 0.8, 0.35                       # BPR353's own factors (dirk_schemes.f90:238-239)
    rad = make_radial(n_r, 64)
    assert np.abs(rad["r"] - h.g.r).max() < 1e-15
    return h, p, rad


def _check(golden, h, row):
    got = np.concatenate([[h.time], h.e_kin()])
    np.testing.assert_allclose(got, golden["reference_out"][row], rtol=RTOL, err_msg=f"row {row}")


def _oracle_host(golden, tweak=None):
    from oracle.oracle import Oracle, Params as OParams
    gs = _sizes(golden)
    o = Oracle(gs["l_max"], minc=4, n_theta=gs["n_theta_max"], n_phi=gs["n_phi_max"], m_max=gs["m_max"],
               threads=min(4, os.cpu_count() or 1))
    h, p, rad = _setup(golden, o.lm2l, o.lm2m)
    op = OParams()
    for n, _ in p._fields_:
        setattr(op, n, getattr(p, n))

    def loop(f):
        op.omega_ic = h.omega_ic
        if tweak:
            tweak(op)
        return o.radial_loop(op, rad, f)
    h.radial_loop = loop
    h._oracle, h._oparams, h._rad = o, op, rad      # for tests that run further oracle batches on the same state (test_testRMSOutputs)
    return h


def test_oracle_radial_loop_reproduces_reference_energies(golden):
    """CPU oracle inside the reference's Runge-Kutta loop: restart state and the first logged row (5 steps, 15 loops);
    the inner core keeps its rotation rate."""
    h = _oracle_host(golden)
    _check(golden, h, 0)
    for _ in range(5):
        h.step()
    _check(golden, h, 1)
    assert abs(h.omega_ic / float(golden["omega_ic1"]) - 1.0) < 1e-8


def test_the_saturated_state_needs_every_term(golden):
    """Negative controls after five steps: no Lorentz force, and an ICB at rest in the loop."""
    for tweak, floor in ((lambda op: setattr(op, "l_mag_LF", 0), 1e-3), (lambda op: setattr(op, "omega_ic", 0.0), 1e-7)):
        h = _oracle_host(golden, tweak)
        for _ in range(5):
            h.step()
        assert np.abs(h.e_kin() / golden["reference_out"][1][1:] - 1.0).max() > floor


@pytest.mark.gpu
def test_gpu_radial_loop_reproduces_reference_energies(golden):
    """The CUDA radial loop inside the reference's Runge-Kutta loop: all five logged rows (25 steps, 75 loops)."""
    from magic_b200 import RadialLoop, Sht
    gs = _sizes(golden)
    s = Sht(gs["l_max"], m_max=gs["m_max"], minc=4, n_theta_max=gs["n_theta_max"], n_phi_max=gs["n_phi_max"])
    h, p, rad = _setup(golden, s.lm2l, s.lm2m)
    rl = RadialLoop(s, p, rad)

    def loop(f):
        rl.set_rotation(0.0, h.omega_ic)
        out = rl.radialLoop(f)
        out["lorentz_torque_ic"], out["lorentz_torque_ma"] = rl.torques()
        return out
    h.radial_loop = loop
    _check(golden, h, 0)
    for row in range(1, len(golden["reference_out"])):
        for _ in range(5):
            h.step()
        _check(golden, h, row)
    assert rl.launch_count() > 0
    rl.finalize()
    s.finalize_sht()
