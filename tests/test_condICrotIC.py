"""End-to-end golden vectors of the reference: samples/dynamo_benchmark_condICrotIC.

The reference's autotest (`samples/dynamo_benchmark_condICrotIC/unitTest.py`, rtol 1e-8) runs the Christensen dynamo
benchmark with a finitely CONDUCTING (sigma_ratio = 1, kbotb = 3) and freely ROTATING (nRotIC = 1) inner core between rigid
walls: l_max = 16, n_r_max = 33 with n_cheb_max = 31, inner core 17 points / 15 even Chebyshev modes, CNAB2 with dt = 1e-4,
e_kin.TAG and e_mag_oc.TAG logged every step.  On the radial-loop side this is the case where
  * lMagNlBc is true (rIter.f90:181-187), so get_nl and the analyses also run on both boundary levels and dVxBhLM carries
    the induction by the moving wall there (get_nl.f90:389-392, get_td.f90:600-617),
  * the ICB level takes v_rigid_boundary with the current omega_ic (nonlinear_bcs.f90:120-175, rIter.f90:570-591),
  * the loop returns the Lorentz torque on the inner core (rIter.f90:279-292, outRot.f90:423-483), which drives omega_ic
    through the z(1,0) torque balance of updateZ.
The inner core spins up from rest to omega_ic = 77 within 60 steps, purely through that torque.  The autotest then RESTARTS
the run from its checkpoint with stress-free walls and l_correct_AMz / AMe (input_restart.nml, 100 more steps): with a
conducting inner core this switches on the nonlinear magnetic boundary condition at the ICB (l_b_nl_icb,
Namelists.f90:713-720), i.e. the loop must also deliver get_br_v_bcs (nonlinear_bcs.f90:24-74, rIter.f90:267-277), and the
inner core is stepped explicitly by the Lorentz torque alone (updateZ.f90:1606-1608).

The Fortran host is restated in numpy (oracle/lmloop.py ShellHost with l_cond_ic / l_rot_ic: coupled outer/inner-core
matrices of get_bMat, even-Chebyshev inner-core grid, z10Mat, finish_exp_tor, finish_exp_mag_ic); the radial loop is the CPU
oracle or the CUDA library through the C ABI, both over all 1000 + 100 steps.  tests/golden/condICrotIC_reference.npz holds
the 1102 rows of reference.out / referenceMag.out (tests/golden/make_condICrotIC_fixture.py).
"""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
RTOL, ATOL = 1e-8, 1e-20          # samples/dynamo_benchmark_condICrotIC/unitTest.py


@pytest.fixture(scope="module")
def golden():
    d = np.load(os.path.join(HERE, "golden", "condICrotIC_reference.npz"))
    return {k: d[k] for k in d.files}


def _sizes(golden):
    from magic_b200.sht import grid_sizes
    gs = grid_sizes(n_phi_tot=int(golden["n_phi_tot"]))
    assert (gs["l_max"], gs["lm_max"]) == (16, 153)
    return gs


def _setup(golden, lm2l, lm2m):
    from magic_b200.workload import make_params, make_radial
    from oracle.lmloop import ShellHost
    n_r = int(golden["n_r_max"])
    kw = {k: float(golden[k]) for k in ("radratio", "ra", "ek", "pr", "prmag", "dtmax", "alpha", "amp_s1", "amp_b1", "sigma_ratio")}
    h = ShellHost(lm2l, lm2m, None, n_r_max=n_r, n_cheb_max=int(golden["n_cheb_max"]), init_s1=int(golden["init_s1"]),
                  init_b1=int(golden["init_b1"]), l_mag=True, l_cond_ic=True, l_rot_ic=True,
                  n_r_ic_max=int(golden["n_r_ic_max"]), n_cheb_ic_max=int(golden["n_cheb_ic_max"]), **kw)
    p = make_params("mhd", n_r, ktopv=int(golden["ktopv"]), kbotv=int(golden["kbotv"]))
    p.l_cond_ic = p.l_rot_ic = 1
    p.ra = kw["ra"]
    p.courfac, p.alffac = float(golden["courfac"]), float(golden["alffac"])
    rad = make_radial(n_r, 16)
    assert np.abs(rad["r"] - h.g.r).max() < 1e-15
    return h, p, rad


def _check(golden, h, row):
    gk = np.concatenate([[h.time], h.e_kin()])
    gm = np.concatenate([[h.time], h.e_mag_oc()])
    np.testing.assert_allclose(gk, golden["e_kin"][row], rtol=RTOL, atol=ATOL, err_msg=f"e_kin row {row}")
    np.testing.assert_allclose(gm, golden["e_mag_oc"][row], rtol=RTOL, atol=ATOL, err_msg=f"e_mag_oc row {row}")


N_FIRST = 1000       # rows 1..1000: input.nml; row 1001: restart state; rows 1002..1101: input_restart.nml
RESTART = dict(ktopv=1, kbotv=1, l_correct_AMz=True, l_correct_AMe=True)


def _oracle_loop(golden, tweak=None):
    from oracle.oracle import Oracle, Params as OParams
    gs = _sizes(golden)
    o = Oracle(gs["l_max"], n_theta=gs["n_theta_max"], n_phi=gs["n_phi_max"], m_max=gs["m_max"], threads=1)
    h, p, rad = _setup(golden, o.lm2l, o.lm2m)
    op = OParams()
    for n, _ in p._fields_:
        setattr(op, n, getattr(p, n))

    def loop(f):
        op.omega_ic = h.omega_ic            # the loop sees the rotation rate of the current state (rIter.f90:570-591)
        if tweak:
            tweak(op)
        return o.radial_loop(op, rad, f)
    h.radial_loop = loop
    return h


def test_start_fields_carry_the_reference_magnetic_energy(golden):
    """Row 0 of referenceMag.out: initB with a conducting inner core (init_fields.f90:1140-1176) differs from the insulating
    start field of samples/dynamo_benchmark; the flow starts at rest."""
    h = _oracle_loop(golden)
    _check(golden, h, 0)


def test_oracle_radial_loop_reproduces_the_first_steps(golden):
    """CPU oracle inside the reference's time loop: the first 30 steps of input.nml, 8 kinetic and 12 magnetic energy columns
    after every step; the inner core is spun up from rest by the Lorentz torque the loop returns.  (All 1000 steps of this
    stage are checked the same way by tests/golden/make_condICrotIC_state.py when it builds the restart state, and by the
    GPU leg below.)"""
    h = _oracle_loop(golden)
    for row in range(1, 31):
        h.step()
        _check(golden, h, row)
    assert 55.0 < h.omega_ic < 65.0 and h.lorentz_torque_ic > 0.0


@pytest.fixture(scope="module")
def first_run(golden):
    """The state after the 1000 steps of input.nml (what checkpoint_end.start holds): tests/golden/condICrotIC_state_1000.npz,
    produced by this host with the CPU oracle, every step checked against reference.out on the way
    (tests/golden/make_condICrotIC_state.py)."""
    d = np.load(os.path.join(HERE, "golden", "condICrotIC_state_1000.npz"))
    return {k: d[k] for k in d.files}


def _restarted(golden, first_run, tweak=None, out_tweak=None):
    from oracle.oracle import Oracle, Params as OParams
    gs = _sizes(golden)
    o = Oracle(gs["l_max"], n_theta=gs["n_theta_max"], n_phi=gs["n_phi_max"], m_max=gs["m_max"], threads=1)
    h, p, rad = _setup(golden, o.lm2l, o.lm2m)
    h.load_state_dict(first_run)
    p.ktopv = p.kbotv = 1
    op = OParams()
    for n, _ in p._fields_:
        setattr(op, n, getattr(p, n))

    def loop(f):
        op.omega_ic = h.omega_ic
        if tweak:
            tweak(op)
        out = o.radial_loop(op, rad, f)
        if out_tweak:
            out_tweak(out)
        return out
    h.radial_loop = loop
    return h


def test_oracle_radial_loop_reproduces_reference_energies(golden, first_run):
    """The restart from the state after 1000 steps: row 1001 is the checkpoint state AFTER startFields has
    applied the angular-momentum correction (before it the axisymmetric toroidal energy is off by 0.32), rows 1002..1101 the
    100 stress-free steps with the nonlinear magnetic boundary condition at the ICB."""
    h = _restarted(golden, first_run)
    assert abs(h.e_kin()[3] / golden["e_kin"][N_FIRST + 1][4] - 1.0) > 0.1
    h.restart(**RESTART)
    _check(golden, h, N_FIRST + 1)
    for row in range(N_FIRST + 2, len(golden["e_kin"])):
        h.step()
        _check(golden, h, row)


def test_the_restart_stage_needs_get_br_v_bcs(golden, first_run):
    """Negative control: without the br*v products of get_br_v_bcs in the ICB boundary condition the energies are off by
    2e-3 (kinetic, axisymmetric toroidal) and 5e-5 (magnetic) after ten steps."""
    h = _restarted(golden, first_run, out_tweak=lambda out: out.__setitem__("br_vp_lm_icb", 0.0 * out["br_vp_lm_icb"]))
    h.restart(**RESTART)
    for _ in range(10):
        h.step()
    row = golden["e_kin"][N_FIRST + 11], golden["e_mag_oc"][N_FIRST + 11]
    assert abs(h.e_kin()[3] / row[0][4] - 1.0) > 1e-4
    assert abs(h.e_mag_oc()[3] / row[1][4] - 1.0) > 1e-6


def test_the_energies_see_the_moving_wall(golden):
    """Negative control: a radial loop that keeps the ICB at rest (omega_ic = 0 in v_rigid_boundary) while the host spins
    the inner core up misses the axisymmetric toroidal energies by 3e-3 (kinetic) and 2e-4 (magnetic) after three steps."""
    h = _oracle_loop(golden, tweak=lambda op: setattr(op, "omega_ic", 0.0))
    for _ in range(3):
        h.step()
    assert abs(h.e_kin()[3] / golden["e_kin"][3][4] - 1.0) > 1e-4
    assert abs(h.e_mag_oc()[3] / golden["e_mag_oc"][3][4] - 1.0) > 1e-5


@pytest.mark.gpu
def test_gpu_radial_loop_reproduces_reference_energies(golden):
    """The CUDA radial loop (magic_rloop_run + magic_rloop_set_rotation + magic_rloop_get_torques, and after the restart
    magic_rloop_get_br_v_bcs) inside the reference's time loop: all 1000 + 100 steps."""
    from magic_b200 import RadialLoop, Sht
    gs = _sizes(golden)
    s = Sht(gs["l_max"], m_max=gs["m_max"], n_theta_max=gs["n_theta_max"], n_phi_max=gs["n_phi_max"])
    h, p, rad = _setup(golden, s.lm2l, s.lm2m)
    rl = RadialLoop(s, p, rad)

    def loop(f):
        rl.set_rotation(0.0, h.omega_ic)
        out = rl.radialLoop(f)
        out["lorentz_torque_ic"], out["lorentz_torque_ma"] = rl.torques()
        return out
    h.radial_loop = loop
    for row in range(1, N_FIRST + 1):
        h.step()
        _check(golden, h, row)
    rl.finalize()
    p.ktopv = p.kbotv = 1
    rl = RadialLoop(s, p, rad)

    def loop2(f):
        rl.set_rotation(0.0, h.omega_ic)
        out = rl.radialLoop(f)
        out["lorentz_torque_ic"], out["lorentz_torque_ma"] = rl.torques()
        out["br_vt_lm_icb"], out["br_vp_lm_icb"] = rl.br_v_bcs("ICB")
        return out
    h.radial_loop = loop2
    h.restart(**RESTART)
    _check(golden, h, N_FIRST + 1)
    for row in range(N_FIRST + 2, len(golden["e_kin"])):
        h.step()
        _check(golden, h, row)
    assert rl.launch_count() > 0
    rl.finalize()
    s.finalize_sht()
