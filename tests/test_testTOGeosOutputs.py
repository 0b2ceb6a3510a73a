"""End-to-end golden vectors of the reference for the torsional-oscillation sums: samples/testTOGeosOutputs (Tay.TAG).

The reference's autotest restarts the saturated benchmark dynamo of samples/boussBenchSat with l_TO on and advances it by 25 BPR353
steps; every five steps outTO (out_TO.f90:213-556) z-averages the (r, theta) arrays that getTO fills inside the radial loop
(rIter.f90:400-404, TO.f90:141-307) on a cylindrical grid and writes Tay.TAG: the energy fractions of the axisymmetric and of the
geostrophic azimuthal flow (from VAS), the Taylorisation measures of the Lorentz stress (dzLFAS), of the Reynolds stress
(dzRstrAS) and of the viscous stress (dzStrAS, getTOfinish), and the kinetic energy.

Host: oracle/lmloop.py DirkShellHost (as tests/test_boussBenchSat.py) and oracle/rms_host.py ToHost (cylmean_otc / cylmean_itc,
simps, the row of Tay.TAG, getTOfinish's viscous stress through toraxi_to_spat).  getTO is the CPU oracle's orc_radial_TO with the
oracle's loop in the time loop (CPU leg: first row) or magic_rloop_to_next / magic_rloop_to through the C ABI with the CUDA loop
(GPU leg: all five rows, called on the steps rIter_cuda_t calls them on; the device's fifteen arrays are also held against the
oracle's, kept fields included).  The autotest also prints seven points of the TO movie with four decimals: single values of VAS,
dzRstrAS, dzAstrAS, dzStrAS, dzLFAS and dzCorAS at given (theta, r) of the first two TO steps (e.g. dzCorAS = 422.0722,
LFfac dzLFAS = -3903.9583), compared here as well -- frame 0 in the CPU leg, frames 0 and 1 in the GPU leg -- and eight points of
TOnhs.TAG / TOshs.TAG (z-averages of VAS, the Reynolds, axisymmetric, Lorentz and viscous stresses, the Taylorisation and the
relative geostrophic flow, which brings in V2AS) spread over the five TO steps (GPU leg; MAGIC_TO_CPU_ROWS=5 runs them on the CPU).  Fixture: tests/golden/testTOGeosOutputs_reference.npz + boussBenchSat_ckpt.npz.
"""
import os

import numpy as np
import pytest

import tests.test_boussBenchSat as bench_sat

HERE = os.path.dirname(os.path.abspath(__file__))
RTOL = 1e-8          # ES16.8 columns, the autotest's tolerance


@pytest.fixture(scope="module")
def golden():
    d = np.load(os.path.join(HERE, "golden", "boussBenchSat_ckpt.npz"))
    g = {k: d[k] for k in d.files}
    r = np.load(os.path.join(HERE, "golden", "testTOGeosOutputs_reference.npz"))
    g.update({k: r[k] for k in r.files})
    return g


def _fields(h):
    return {k: np.ascontiguousarray(v) for k, v in h.fields_Rloc().items()}


def _check_movie_points(golden, h, to_host, arrays, frame):
    """The TO movie stores VAS, dzRstrAS, dzAstrAS, dzStrAS, LFfac dzLFAS, dzCorAS (and dzdVpAS) of every TO step in single
    precision; the autotest prints seven of its points with four decimals (unitTest.py:50-53)."""
    LFfac = 1.0 / (float(golden["ek"]) * float(golden["prmag"]))
    get = {0: lambda t, r: arrays[r, 1, t], 1: lambda t, r: arrays[r, 3, t], 2: lambda t, r: arrays[r, 4, t],
           3: lambda t, r: to_host.dzStrAS()[t, r], 4: lambda t, r: LFfac * arrays[r, 5, t], 5: lambda t, r: arrays[r, 2, t]}
    for q, ((fr, t, r), ref) in enumerate(zip(golden["movie_points"], golden["movie_values"])):
        if fr != frame or q not in get:
            continue
        got = float(np.float32(get[q](t, r)))
        assert abs(got - ref) <= 5.1e-5 + 1e-7 * abs(ref), (q, got, ref)


def _check_hemi_points(golden, means, frame):
    """Eight points of TOnhs.TAG / TOshs.TAG (z-averages on the cylindrical grid, single precision, four printed decimals;
    unitTest.py:56-63).  `frame` counts the TO steps from 1: entry [k, n_s] of the readers is TO step k + 1."""
    LFfac = 1.0 / (float(golden["ek"]) * float(golden["prmag"]))
    north = [("Vp", 2, 18, 1.0), ("Rstr", 3, 11, 1.0), ("Astr", 0, 30, 1.0), ("LF", 2, 21, LFfac)]
    south = [(None, 3, 12, 1.0), ("Str", 1, 9, 1.0), ("Tay", 4, 27, 1.0), ("VpR", 2, 21, 1.0)]      # dvp: host-only, ~ 0
    for hemi, pts, ref in ((0, north, golden["nhs_values"]), (1, south, golden["shs_values"])):
        for (name, k, n_s, fac), r in zip(pts, ref):
            if name is None or k + 1 != frame:
                continue
            got = float(np.float32(fac * means[name][hemi][n_s]))
            assert abs(got - r) <= 5.1e-5 + 1e-7 * abs(r), (name, hemi, got, r)


def _run(golden, h, to_host, to_next, to, n_rows):
    """step_time.f90:355-382 with n_TO_step = 5: getTOnext's kept fields at the first stage of steps 5, 10, ..., getTO (with the
    previous time step as dtLast) and outTO at the first stage of steps 6, 11, ..."""
    n_to = int(golden["n_TO_step"])
    dt = float(golden["dt"][0])
    rows = []
    for step in range(1, n_rows * n_to + 2):        # `step` = n_time_step of the reference; its first stage sees `step - 1` steps done
        f = _fields(h)
        if step > 2 and (step - 1) % n_to == 0:
            arrays = to(f, dt)
            means = to_host.cyl_means(arrays, with_hemi_files=True)
            rows.append(to_host.row(arrays, h.e_kin(), means))
            _check_hemi_points(golden, means, frame=len(rows))
            np.testing.assert_allclose(rows[-1], golden["Tay"][len(rows) - 1], rtol=RTOL, err_msg=f"Tay row {len(rows) - 1}")
            _check_movie_points(golden, h, to_host, arrays, frame=len(rows) - 1)
            if len(rows) == n_rows:
                break
        if step % n_to == 0:
            to_next(f)
        h.step()
    return np.array(rows)


def test_oracle_getTO_reproduces_Tay(golden):
    from oracle.rms_host import ToHost
    h = bench_sat._oracle_host(golden)
    o, op, rad = h._oracle, h._oparams, h._rad
    state = {"last": None}

    def with_omega():
        op.omega_ic = h.omega_ic
        return op
    rows = _run(golden, h, ToHost(h, o.theta_ord, o.toraxi_to_spat),
                to_next=lambda f: state.update(last=o.radial_TO(with_omega(), rad, f, 0)),
                to=lambda f, dt: o.radial_TO(with_omega(), rad, f, 1, dtLast=dt, last=state["last"]),
                n_rows=int(os.environ.get("MAGIC_TO_CPU_ROWS", "1")))
    assert rows.shape[1] == 7
    # negative control: the Taylorisation of the Lorentz stress needs dzLFAS -- with the field halved it is unchanged (a ratio),
    # with br and bt of one sign flipped ... simpler: the Reynolds measure collapses to 1 for an axisymmetric flow
    f = _fields(h)
    fa = {k: np.where(h.lm2m[None, :] == 0, v, 0.0) for k, v in f.items()}
    row = ToHost(h, o.theta_ord, o.toraxi_to_spat).row(o.radial_TO(with_omega(), rad, fa, 1, dtLast=1.0), h.e_kin())
    assert abs(row[4] - golden["Tay"][0, 4]) > 1e-3


@pytest.mark.gpu
def test_gpu_getTO_reproduces_Tay(golden):
    """magic_rloop_to_next / magic_rloop_to with the CUDA radial loop in the time loop: all five rows of Tay.start; at every row the
    fifteen (r, theta) arrays of the device agree with the oracle's evaluated on the same fields and the same kept fields."""
    from magic_b200 import RadialLoop, Sht
    from oracle.rms_host import ToHost
    gs = bench_sat._sizes(golden)
    s = Sht(gs["l_max"], m_max=gs["m_max"], minc=4, n_theta_max=gs["n_theta_max"], n_phi_max=gs["n_phi_max"])
    h, p, rad = bench_sat._setup(golden, s.lm2l, s.lm2m)
    rl = RadialLoop(s, p, rad)
    ho = bench_sat._oracle_host(golden)                 # only for its oracle handle and parameters
    o, op = ho._oracle, ho._oparams
    state = {"last": None}

    def loop(f):
        rl.set_rotation(0.0, h.omega_ic)
        out = rl.radialLoop(f)
        out["lorentz_torque_ic"], out["lorentz_torque_ma"] = rl.torques()
        return out
    h.radial_loop = loop

    def to_next(f):
        op.omega_ic = h.omega_ic
        state["last"] = o.radial_TO(op, rad, f, 0)
        rl.to_next(f)

    def to(f, dt):
        rl.set_rotation(0.0, h.omega_ic)
        got = rl.to(f, dt)
        op.omega_ic = h.omega_ic
        ref = o.radial_TO(op, rad, f, 1, dtLast=dt, last=state["last"])
        for q in range(ref.shape[1]):
            scale = np.abs(ref[:, q]).max()
            # the four time-derivative arrays are differences of nearly equal fields (a steadily drifting dynamo) over dt = 2e-4
            tol = 1e-9 if 10 <= q <= 13 else 1e-11
            assert np.abs(got[:, q] - ref[:, q]).max() <= tol * scale, q
        return got
    theta_ord, _ = s.get_grid()
    rows = _run(golden, h, ToHost(h, theta_ord, lambda tl, lcut: s.toraxi_to_spat(tl, lcut)), to_next, to, n_rows=len(golden["Tay"]))
    assert rows.shape == (5, 7)
    rl.finalize()
    s.finalize_sht()
