"""GPU parity: the batched radial loop (rIter.f90:94-464) through the C ABI vs the CPU oracle."""
import numpy as np
import pytest

from tests.util import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-12


def oracle_params(p):
    from oracle.oracle import Params as OParams
    op = OParams()
    for n, _ in p._fields_:
        setattr(op, n, getattr(p, n))
    return op


def run_both(l_max, n_r_max, physics, levels, level_chunk=0, ktopv=2, kbotv=2, minc=1, n_phi_tot=0, l_R=None, seed=3,
             full_sphere=False, tweak=None):
    from magic_b200 import RadialLoop, Sht, grid_sizes
    from magic_b200.workload import make_fields, make_params, make_radial
    from oracle.oracle import Oracle
    gs = grid_sizes(l_max=l_max, n_phi_tot=n_phi_tot, minc=minc)
    o = Oracle(gs["l_max"], minc=minc, n_theta=gs["n_theta_max"], n_phi=gs["n_phi_max"], m_max=gs["m_max"])
    s = Sht(gs["l_max"], m_max=gs["m_max"], minc=minc, n_theta_max=gs["n_theta_max"], n_phi_max=gs["n_phi_max"])
    p = make_params(physics, n_r_max, ktopv=ktopv, kbotv=kbotv)
    p.l_full_sphere = 1 if full_sphere else 0
    if tweak:
        for k, v in tweak.items():
            setattr(p, k, v)
    rad_full = make_radial(n_r_max, gs["l_max"], l_R=l_R, anel=(physics == "anel"))
    idx = np.array(levels) - 1
    rad = {k: np.ascontiguousarray(v[idx]) for k, v in rad_full.items()}
    fields = make_fields(physics, o.lm2l, o.lm2m, len(levels), seed)
    if tweak and tweak.get("l_phase_field"):   # a phase field around 1/2 with structure at every degree
        fields["phi"] = 0.05 * make_fields("hydro", o.lm2l, o.lm2m, len(levels), seed + 1)["w"]
        fields["phi"][:, 0] = 0.5 * np.sqrt(4.0 * np.pi)
    rl = RadialLoop(s, p, rad, level_chunk=level_chunk)
    got = rl.radialLoop(fields)
    got["lorentz_torque_ic"], got["lorentz_torque_ma"] = rl.torques()
    ref = o.radial_loop(oracle_params(p), rad, fields)
    # Conditioning probe: the oracle's response to a last-bits (1e-15 relative) perturbation of its inputs.
    # Outputs that are small differences of large sums (e.g. the toroidal part of the nonlinear force,
    # |T| ~ 1e-3 |S| for these rough synthetic fields) cannot agree better than this between ANY two
    # correctly rounded evaluation orders -- the reference's own -O2 and -O3 -mfma builds differ by more.
    prng = np.random.default_rng(12345)
    pert = {k: v * (1.0 + 1e-15 * prng.standard_normal(v.shape)) for k, v in fields.items()}
    noise = o.radial_loop(oracle_params(p), rad, pert)
    timing = rl.last_timing()
    rl.finalize()
    s.finalize_sht()
    return o, p, rad, got, ref, (noise, timing)


def compare(o, p, rad, got, ref, extra, names):
    """rel. L2 <= 1e-12, or 10x the oracle's own sensitivity to a 1e-15 input perturbation where that is larger."""
    noise = extra[0]
    nR = rad["nR"]
    bulk = (nR != 1) & (nR != p.n_r_max)
    for nm in names:
        g, r = got[nm], ref[nm]
        if nm in ("dVxBhLM", "dVxVhLM", "dVSrLM", "dVXirLM"):
            sel = slice(None)  # defined on every level (get_td.f90:299-307,511-519,600-617)
        else:
            sel = bulk         # d?dt are only written on bulk levels (get_td.f90:153,330,400,478)
        lo = 1 if nm == "dpdt" else 0  # dpdt(lm=1) is never written by the reference
        floor = rel_l2(noise[nm][sel][:, lo:], r[sel][:, lo:])
        assert rel_l2(g[sel][:, lo:], r[sel][:, lo:]) < max(TOL, 10.0 * floor), (nm, floor)
    assert np.allclose(got["dtrkc"], ref["dtrkc"], rtol=1e-12)
    assert np.allclose(got["dthkc"], ref["dthkc"], rtol=1e-12)


MHD_OUT = ["dwdt", "dzdt", "dpdt", "dsdt", "dbdt", "djdt", "dVxBhLM", "dVSrLM"]
HYDRO_OUT = ["dwdt", "dzdt", "dpdt", "dsdt", "dVSrLM"]


def test_mhd_l16_with_rigid_boundaries():
    """dynamo_benchmark shape: l_max=16, n_r=33, rigid walls; levels include both boundaries."""
    o, p, rad, got, ref, ex = run_both(16, 33, "mhd", [1, 2, 3, 17, 31, 32, 33])
    compare(o, p, rad, got, ref, ex, MHD_OUT)


def test_mhd_l16_chunked_equals_unchunked():
    a = run_both(16, 33, "mhd", list(range(1, 12)), level_chunk=0)
    b = run_both(16, 33, "mhd", list(range(1, 12)), level_chunk=4)  # chunks of 4,4,3 -> two layouts
    for nm in MHD_OUT:
        assert np.array_equal(a[3][nm], b[3][nm]), nm
    compare(*b, MHD_OUT)


def test_mhd_stress_free_boundaries():
    """ktopv=kbotv=1: lMagNlBc true, get_nl runs on the boundary with nBc=1 (rIter.f90:181-187, get_nl.f90:389-392)."""
    o, p, rad, got, ref, ex = run_both(16, 33, "mhd", [1, 2, 16, 33], ktopv=1, kbotv=1)
    compare(o, p, rad, got, ref, ex, MHD_OUT)


def test_hydro_minc3_variable_lcut():
    """full_sphere-like truncation (minc=3, n_phi_tot=96) with l_R varying per level (radial.f90:291-307)."""
    l_R = np.full(12, 32)
    l_R[6:] = [30, 27, 23, 18, 12, 5]
    o, p, rad, got, ref, ex = run_both(0, 12, "hydro", list(range(1, 13)), minc=3, n_phi_tot=96, l_R=l_R)
    compare(o, p, rad, got, ref, ex, HYDRO_OUT)
    for i, lc in enumerate(rad["l_R"]):
        if rad["nR"][i] in (1, 12):
            continue
        assert np.all(got["dsdt"][i][(o.lm2l > lc)] == 0)


def test_anelastic_hydro_ugradu():
    """hydro_bench_anel shape (l_adv_curl=.false., viscous heating; get_nl.f90:274-308,402-436)."""
    o, p, rad, got, ref, ex = run_both(0, 9, "anel", list(range(1, 10)), n_phi_tot=96, ktopv=1, kbotv=1)
    compare(o, p, rad, got, ref, ex, HYDRO_OUT)


def test_mhd_l96_bulk():
    o, p, rad, got, ref, ex = run_both(0, 97, "mhd", [40, 41, 42], n_phi_tot=288)
    compare(o, p, rad, got, ref, ex, MHD_OUT)


def test_run_to_run_bitwise():
    a = run_both(16, 33, "mhd", [5, 6, 7, 8])
    b = run_both(16, 33, "mhd", [5, 6, 7, 8])
    for nm in MHD_OUT + ["dtrkc", "dthkc"]:
        assert np.array_equal(a[3][nm], b[3][nm]), nm


@pytest.mark.parametrize("physics,minc,n_phi_tot", [("hydro", 1, 0), ("mhd", 1, 0), ("hydro", 3, 96)])
def test_full_sphere_centre(physics, minc, n_phi_tot):
    """samples/full_sphere geometry: the r=0 level takes v_center_sphere (nonlinear_bcs.f90:177-224, l=1 modes of ddw / ddb)
    and skips the Courant check (rIter.f90:295); l_R varies with radius (radial.f90:291-307)."""
    n_r = 10
    l_max = 0 if n_phi_tot else 16
    lm = 32 if n_phi_tot else 16
    l_R = np.minimum(lm, (1 + lm * np.sqrt(np.linspace(1.0, 0.05, n_r) / 0.4)).astype(int))
    o, p, rad, got, ref, ex = run_both(l_max, n_r, physics, list(range(1, n_r + 1)), minc=minc, n_phi_tot=n_phi_tot, l_R=l_R,
                                       ktopv=1, kbotv=1, full_sphere=True)
    compare(o, p, rad, got, ref, ex, MHD_OUT if physics == "mhd" else HYDRO_OUT)
    assert got["dtrkc"][-1] == 1e10 and got["dthkc"][-1] == 1e10


DC_OUT = ["dwdt", "dzdt", "dsdt", "dVSrLM", "dVxVhLM"]


@pytest.mark.parametrize("physics,full_sphere", [("hydro", True), ("hydro", False), ("mhd", False)])
def test_double_curl_form(physics, full_sphere):
    """l_double_curl (forced by radial_scheme='FD', Namelists.f90:299-304; what samples/full_sphere runs): get_dwdt_double_curl
    (get_td.f90:199-309) with its Coriolis couplings cut at l_R(nR), and the horizontal advection handed out as dVxVhLM for
    finish_exp_pol (updateWP.f90:1002-1031); no dpdt.  The oracle's branch is pinned to the reference by
    tests/test_full_sphere.py."""
    n_r = 10
    mhd = physics == "mhd"
    l_R = None if mhd else np.minimum(32, (1 + 32 * np.sqrt(np.linspace(1.0, 0.05, n_r) / 0.4)).astype(int))
    o, p, rad, got, ref, ex = run_both(16 if mhd else 0, n_r, physics, list(range(1, n_r + 1)), minc=1 if mhd else 3,
                                       n_phi_tot=0 if mhd else 96, l_R=l_R, ktopv=1, kbotv=1, full_sphere=full_sphere,
                                       tweak=dict(l_double_curl=1))
    assert np.linalg.norm(ref["dVxVhLM"]) > 0 and np.linalg.norm(ref["dpdt"]) == 0
    compare(o, p, rad, got, ref, ex, DC_OUT + (["dbdt", "djdt", "dVxBhLM"] if mhd else []))
    assert np.all(got["dpdt"] == 0)


def test_hydro_bench_anel_shape():
    """BASELINE config 2 (samples/hydro_bench_anel as shipped: n_phi_tot=288 -> l_max=96, n_r=97, anelastic hydro with
    u.grad u advection and viscous heating, stress-free walls): a CMB level, three bulk levels and the ICB level."""
    o, p, rad, got, ref, ex = run_both(0, 97, "anel", [1, 2, 48, 96, 97], n_phi_tot=288, ktopv=1, kbotv=1)
    assert o.l_max == 96 and o.lm_max == 4753
    compare(o, p, rad, got, ref, ex, HYDRO_OUT)


def test_stress_free_conducting_walls_get_br_v_bcs():
    """l_b_nl_cmb / l_b_nl_icb (stress-free walls + conducting mantle / inner core, Namelists.f90:713-729): the boundary
    levels also deliver get_br_v_bcs (nonlinear_bcs.f90:24-74, rIter.f90:267-277), the products the magnetic boundary
    conditions of updateB use.  Rotating walls so that the omega * sin^2(theta) term is exercised."""
    tw = dict(l_cond_ic=1, l_cond_ma=1, l_rot_ic=1, l_rot_ma=1, omega_ic=0.37, omega_ma=-0.21)
    from magic_b200 import RadialLoop, Sht
    from magic_b200.workload import make_fields, make_params, make_radial
    from oracle.oracle import Oracle
    o = Oracle(16)
    s = Sht(16)
    p = make_params("mhd", 33, ktopv=1, kbotv=1)
    for k, v in tw.items():
        setattr(p, k, v)
    levels = np.array([1, 2, 17, 32, 33])
    rad = {k: np.ascontiguousarray(v[levels - 1]) for k, v in make_radial(33, 16).items()}
    fields = make_fields("mhd", o.lm2l, o.lm2m, len(levels), 11)
    rl = RadialLoop(s, p, rad)
    got = rl.radialLoop(fields)
    ref = o.radial_loop(oracle_params(p), rad, fields)
    for bc in ("cmb", "icb"):
        vt, vp = rl.br_v_bcs(bc.upper())
        assert np.linalg.norm(ref["br_vt_lm_" + bc]) > 0 and np.linalg.norm(ref["br_vp_lm_" + bc]) > 0
        assert rel_l2(vt, ref["br_vt_lm_" + bc]) < TOL and rel_l2(vp, ref["br_vp_lm_" + bc]) < TOL, bc
    for nm in ("dbdt", "djdt", "dVxBhLM"):
        assert rel_l2(got[nm][1:4], ref[nm][1:4]) < 1e-11, nm
    # a loop without the boundary level, or without the physics, has no such products
    rl2 = RadialLoop(s, make_params("mhd", 33), rad)
    from magic_b200 import MagicError
    with pytest.raises(MagicError, match="no nonlinear magnetic boundary"):
        rl2.br_v_bcs("CMB")
    rl.finalize()
    rl2.finalize()
    s.finalize_sht()


def test_conducting_rotating_walls_and_lorentz_torques():
    """boussBenchSat physics (conducting, rotating inner core; here also the mantle): rigid walls move with omega
    (v_rigid_boundary), get_nl runs on the boundary levels (lMagNlBc) and the Lorentz torques are the quadrature of
    Br*Bp over the wall (rIter.f90:279-292,461, outRot.f90:423-483)."""
    tw = dict(l_cond_ic=1, l_rot_ic=1, l_cond_ma=1, l_rot_ma=1, omega_ic=0.37, omega_ma=-0.21)
    o, p, rad, got, ref, ex = run_both(16, 33, "mhd", [1, 2, 3, 31, 32, 33], tweak=tw)
    compare(o, p, rad, got, ref, ex, MHD_OUT)
    assert ref["lorentz_torque_ic"] != 0 and ref["lorentz_torque_ma"] != 0
    assert abs(got["lorentz_torque_ic"] / ref["lorentz_torque_ic"] - 1) < 1e-11
    assert abs(got["lorentz_torque_ma"] / ref["lorentz_torque_ma"] - 1) < 1e-11
    # rotation rates can change from step to step
    assert np.linalg.norm(got["dVxBhLM"][0]) > 0  # moving wall drags field lines: non-zero boundary induction term


def test_phase_field_branch():
    """l_phase_field (get_nl.f90:333-344, rIter.f90:509,698): the phase field is synthesised, penalises the velocity in the
    advection terms and yields dphidt = scal_to_SH(phiTerms) on bulk levels; Boussinesq hydro with stress-free walls."""
    tw = dict(l_phase_field=1, epsPhase=0.03, phaseDiffFac=1.0, penaltyFac=0.5, tmelt=0.11)
    o, p, rad, got, ref, ex = run_both(21, 17, "hydro", [1, 2, 3, 9, 16, 17], ktopv=1, kbotv=1, tweak=tw)
    compare(o, p, rad, got, ref, ex, HYDRO_OUT + ["dphidt"])
    o2, p2, rad2, got2, ref2, ex2 = run_both(21, 17, "hydro", [1, 2, 3, 9, 16, 17], ktopv=1, kbotv=1)
    bulk = (rad["nR"] != 1) & (rad["nR"] != 17)
    assert rel_l2(got["dwdt"][bulk], got2["dwdt"][bulk]) > 1e-3, "the penalty term is not seen"


def test_phase_field_with_mhd_and_u_grad_u():
    """The same branch next to the magnetic terms and the u.grad u advection form (EXTRA kernel variant with MAG)."""
    tw = dict(l_phase_field=1, epsPhase=0.05, phaseDiffFac=2.0, penaltyFac=1.0, tmelt=0.3, l_adv_curl=0)
    o, p, rad, got, ref, ex = run_both(16, 9, "mhd", list(range(1, 10)), tweak=tw)
    compare(o, p, rad, got, ref, ex, MHD_OUT + ["dphidt"])
