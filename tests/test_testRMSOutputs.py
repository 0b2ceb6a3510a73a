"""End-to-end golden vectors of the reference for the r.m.s. force balance and the dynamo terms: samples/testRMSOutputs (first stage).

The reference's autotest restarts the saturated benchmark dynamo of samples/boussBenchSat (conducting, freely rotating inner
core; l_max = 64, minc = 4, n_r_max = 33) with l_RMS on and advances it by 50 steps of the IMEX Runge-Kutta scheme BPR353, logging
dtVrms.TAG every 10 steps: sixteen columns -- inertia, Coriolis, Lorentz, advection, viscous, buoyancy, pressure-gradient r.m.s.
forces and seven force-balance ratios.  All but the viscous column are built on the fourteen spectra the radial loop returns on
lRmsCalc steps (get_nl with every level as bulk, get_nl_RMS, transform_to_lm_RMS: rIter.f90:215-252, 710; RMS.f90:469-610).
On the same steps l_RMS switches on get_dtBLM (step_time.f90:386, rIter.f90:388-391) and dtBrms.TAG is written: eleven columns, of
which the dynamo terms PdynRms, TdynRms, the omega-effect ratios and the dipole parts are built on the eleven spectra of the
get_dtBLM batch (magic_rloop_dtb); the others are the time derivative and the diffusion of the field (host only).

Host: oracle/lmloop.py DirkShellHost (as in tests/test_boussBenchSat.py; it gained the l = 0 pressure solve of updateWP.f90:358-394,
which only this diagnostic reads) and oracle/rms_host.py (compute_lm_forces, init_rNB, get_force, the row of dtVrms; get_dH_dtBLM,
get_dtBLMfinish, get_PolTorRms, the row of dtBrms).  The batches are the CPU oracle's orc_radial_RMS / orc_radial_dtB with the oracle's loop in the time loop (CPU leg: first row) or magic_rloop_rms_keep /
magic_rloop_rms / magic_rloop_dtb through the C ABI with the CUDA loop (GPU leg: all five rows of both files, called as rIter_cuda_t
calls them -- keep at the first stage of every step, the batches on the logged steps).  The pressure-gradient column moves in its sixth digit between the
first and the later rows (the l = 0 pressure sees the explicit term of the stage that solved it), which both legs reproduce.
Fixture: tests/golden/testRMSOutputs_reference.npz (tests/golden/make_testRMSOutputs_fixture.py) + boussBenchSat_ckpt.npz.
"""
import os

import numpy as np
import pytest

import tests.test_boussBenchSat as bench_sat

HERE = os.path.dirname(os.path.abspath(__file__))
RTOL_FORCES = 1e-8      # ES16.8 columns, the autotest's tolerance
RTOL_RATIOS = 5e-7      # ES14.6 columns: seven printed digits


@pytest.fixture(scope="module")
def golden():
    d = np.load(os.path.join(HERE, "golden", "boussBenchSat_ckpt.npz"))
    g = {k: d[k] for k in d.files}
    r = np.load(os.path.join(HERE, "golden", "testRMSOutputs_reference.npz"))
    g.update({k: r[k] for k in r.files})
    return g


def _check(golden, row, got):
    ref = golden["dtVrms"][row]
    np.testing.assert_allclose(got[:9], ref[:9], rtol=RTOL_FORCES, atol=1e-30, err_msg=f"dtVrms row {row}: forces")
    np.testing.assert_allclose(got[9:], ref[9:], rtol=RTOL_RATIOS, err_msg=f"dtVrms row {row}: balances")


def _fields(h):
    f = {k: np.ascontiguousarray(v) for k, v in h.fields_Rloc().items()}
    f["p"] = np.ascontiguousarray(h.p)                     # transform_to_grid_RMS reads the pressure (RMS.f90:562-574)
    return f


def _check_dtb(golden, row, got):
    np.testing.assert_allclose(got, golden["dtBrms"][row], rtol=RTOL_FORCES, err_msg=f"dtBrms row {row}")


def _run(golden, h, rms_host, keep, batch, n_rows, dtb_batch=None):
    """The reference's sequence: at the first stage of every step the previous velocity is refreshed (get_nl_RMS, RMS.f90:545-551);
    on the logged steps the batch runs first and dtVrms follows (step_time.f90:384, output.f90)."""
    from oracle.rms_host import DtbHost
    n_log = int(golden["n_log_step"])
    rows, dtb_host, start = [], DtbHost(h), None
    for step in range(n_rows * n_log + 1):
        f = _fields(h)
        if step and step % n_log == 0:
            rows.append(rms_host.row(batch(f), CorFac=1.0 / float(golden["ek"])))
            _check(golden, len(rows) - 1, rows[-1])
            if dtb_batch is not None:   # get_dtBLM of the same step, then dtBrms (output.f90:508-531)
                _check_dtb(golden, len(rows) - 1, dtb_host.row(dtb_batch(f), start[0], start[1], float(golden["dt"][0])))
            if len(rows) == n_rows:
                break
        keep(f)
        start = (h.b.copy(), h.aj.copy())     # the field the time derivative of updateB.f90:1638-1643 is taken against
        h.step()
    return np.array(rows)


def test_oracle_rms_batch_reproduces_dtVrms(golden):
    from oracle.rms_host import RmsHost
    h = bench_sat._oracle_host(golden)
    o, op, rad = h._oracle, h._oparams, h._rad
    state = {}
    rows = _run(golden, h, RmsHost(h, rCut=float(golden["rCut"]), rDea=float(golden["rDea"])),
                keep=lambda f: state.update(old={k: f[k].copy() for k in ("w", "dw", "z")}),
                batch=lambda f: o.radial_RMS(_with_omega(op, h), rad, f, state["old"], float(golden["dt"][0])), n_rows=1,
                dtb_batch=lambda f: o.radial_dtB(_with_omega(op, h), rad, f))
    # negative controls: without the curl-form correction of the pressure term, or with the present velocity as the "previous"
    # one (no inertia), the golden row is missed
    f = _fields(h)
    rh = RmsHost(h)
    bad = o.radial_RMS(_with_omega(op, h), rad, f, {k: f[k] for k in ("w", "dw", "z")}, float(golden["dt"][0]))
    assert abs(rh.row(bad, 1.0 / float(golden["ek"]))[1] / golden["dtVrms"][0, 1] - 1.0) > 1e-3          # InerRms
    rq = o.radial_RMS(_with_omega(op, h), rad, f, state["old"], float(golden["dt"][0]))
    rq[3] = 0.0                                                                                          # dpkindrLM
    assert abs(rh.row(rq, 1.0 / float(golden["ek"]))[8] / golden["dtVrms"][0, 8] - 1.0) > 1e-4           # PreRms
    assert rows.shape == (1, 16)
    # ... and the dynamo terms need the products: with BtVr and BrVt swapped PdynRms only changes sign inside, but without the
    # omega-effect products the ratio columns vanish
    from oracle.rms_host import DtbHost
    dtb = o.radial_dtB(_with_omega(op, h), rad, f)
    dtb[8:10] = 0.0
    assert DtbHost(h).row(dtb, h.b, h.aj, 1.0)[7] == 0.0


def _with_omega(op, h):
    op.omega_ic = h.omega_ic
    return op


@pytest.mark.gpu
def test_gpu_rms_batch_reproduces_dtVrms(golden):
    """magic_rloop_rms_keep / magic_rloop_rms with the CUDA radial loop in the time loop: all five rows of dtVrms.start."""
    from magic_b200 import RadialLoop, Sht
    from oracle.rms_host import RmsHost
    gs = bench_sat._sizes(golden)
    s = Sht(gs["l_max"], m_max=gs["m_max"], minc=4, n_theta_max=gs["n_theta_max"], n_phi_max=gs["n_phi_max"])
    h, p, rad = bench_sat._setup(golden, s.lm2l, s.lm2m)
    rl = RadialLoop(s, p, rad)

    def loop(f):
        rl.set_rotation(0.0, h.omega_ic)
        out = rl.radialLoop(f)
        out["lorentz_torque_ic"], out["lorentz_torque_ma"] = rl.torques()
        return out
    h.radial_loop = loop
    rows = _run(golden, h, RmsHost(h, rCut=float(golden["rCut"]), rDea=float(golden["rDea"])), keep=rl.rms_keep,
                batch=lambda f: rl.rms(f, float(golden["dt"][0])), n_rows=len(golden["dtVrms"]),
                dtb_batch=lambda f: (rl.set_rotation(0.0, h.omega_ic), rl.dtb(f))[1])    # the wall values of get_dtBLM see the present omega_ic
    assert rows.shape == (5, 16)
    assert abs(rows[1, 8] / rows[0, 8] - 1.0) > 5e-6        # the pressure-gradient column does move after the first row
    rl.finalize()
    s.finalize_sht()
