"""In-loop diagnostics of log steps (rIter.f90:303-373): get_helicity, get_hemi (outMisc.f90:991-1167),
get_visc_heat (power.f90:384-441), get_perpPar, get_fluxes, get_nlBLayers (outPar.f90:470-726).

CPU: the oracle's restatement (oracle/magic_oracle_diag.inc) against closed forms that do not go through the grid --
Parseval for the hemispheric energies, solid-body rotation for helicity / viscous heating / E_perp.
GPU (-m gpu): magic_rloop_diagnostics through the C ABI against the oracle on the same seeded spectra: Boussinesq MHD with
rigid walls, anelastic hydro with stress-free walls, rotating rigid walls, lRmsCalc (boundaries as bulk), device pointers.
The golden pin of these routines (helicity.TAG / hemi.TAG / power.TAG of samples/testOutputs) is tests/test_testOutputs.py.
"""
import numpy as np
import pytest

from magic_b200.riter import (DIAG_FLUX, DIAG_HEL, DIAG_HEMI, DIAG_PERPPAR, DIAG_POWER, DIAG_RMSBULK, DIAG_VISCBC, NDIAG)
from magic_b200.workload import make_fields, make_params, make_radial

ALL = DIAG_HEL | DIAG_HEMI | DIAG_POWER | DIAG_PERPPAR | DIAG_FLUX | DIAG_VISCBC
TOL = 1e-12   # relative to the largest entry of a slot over the levels (sums of ~1e4 positive and negative terms)


def _oracle(l_max, minc=1):
    from oracle.oracle import Oracle
    return Oracle(l_max, minc=minc)


def _oparams(p):
    from oracle.oracle import Params as OParams
    op = OParams()
    for n, _ in p._fields_:
        setattr(op, n, getattr(p, n))
    return op


def _case(physics, l_max, n_r_max, lm2l, lm2m, seed, ktopv=2, kbotv=2, anel=False):
    p = make_params(physics, n_r_max, ktopv=ktopv, kbotv=kbotv)
    rad = make_radial(n_r_max, l_max, anel=anel)
    f = make_fields(physics, lm2l, lm2m, n_r_max, seed)
    rng = np.random.default_rng(seed + 7)
    for nm, src in (("p", "s"), ("ds", "s")):   # pressure and the radial entropy derivative: any smooth spectra will do
        f[nm] = f[src] * (0.3 + rng.random()) + 0.1 * f["w"]
    return p, rad, f


def test_oracle_hemispheric_energies_obey_parseval():
    """get_hemi against the spectral energy: north + south of 1/2 int (vr^2/r^2 + (vt^2 + vp^2)/sin^2) dOmega (the grid fields
    are r^2 u_r and r sin(theta) u_h) equals 1/2 sum_lm (2 - delta_m0) [ l^2(l+1)^2/r^2 |w|^2 + l(l+1) (|dw|^2 + |z|^2) ] for
    an orthonormal basis."""
    l_max, n_r = 16, 5
    o = _oracle(l_max)
    p, rad, f = _case("mhd", l_max, n_r, o.lm2l, o.lm2m, 3)
    d = o.radial_diagnostics(_oparams(p), rad, f, DIAG_HEMI | DIAG_RMSBULK)
    l, m = o.lm2l.astype(float), o.lm2m
    fac = np.where(m == 0, 1.0, 2.0)
    dLh = l * (l + 1.0)
    for i in range(n_r):
        or2 = rad["or2"][i]
        for (q, s, t, n0) in (("w", "dw", "z", 9), ("b", "db", "aj", 13)):
            e = 0.5 * np.sum(fac * (dLh ** 2 * or2 * abs(f[q][i]) ** 2 + dLh * (abs(f[s][i]) ** 2 + abs(f[t][i]) ** 2)))
            assert abs(d[i, n0] + d[i, n0 + 1] - e) < 1e-12 * e


def test_oracle_solid_body_rotation_closed_forms():
    """A rigid rotation about z (toroidal l=1, m=0: z = c r^2) has no helicity, no viscous heating, and all its kinetic
    energy perpendicular to the axis: E_perp = E_perp_axi = the hemi energies (up to the 2 pi / r^2 factors of the
    routines), E_par = 0."""
    l_max, n_r = 8, 4
    o = _oracle(l_max)
    p = make_params("hydro", n_r)
    rad = make_radial(n_r, l_max)
    lm10 = int(np.where((o.lm2l == 1) & (o.lm2m == 0))[0][0])
    f = {k: np.zeros((n_r, o.lm_max), dtype=complex) for k in ("w", "dw", "ddw", "z", "dz", "s", "ds", "p")}
    c = 0.7
    f["z"][:, lm10] = c * rad["r"] ** 2
    f["dz"][:, lm10] = 2 * c * rad["r"]
    d = o.radial_diagnostics(_oparams(p), rad, f, ALL | DIAG_RMSBULK)
    scale = np.abs(d).max()
    assert np.abs(d[:, 0:9]).max() < 1e-13 * scale          # helicity
    assert np.abs(d[:, 17]).max() < 1e-10 * scale           # viscous heating of a rigid rotation
    assert np.abs(d[:, 19]).max() < 1e-13 * scale and np.abs(d[:, 21]).max() < 1e-13 * scale   # E_par
    np.testing.assert_allclose(d[:, 18], d[:, 20], rtol=1e-12)                                    # axisymmetric flow
    np.testing.assert_allclose(2 * np.pi * d[:, 18], (d[:, 9] + d[:, 10]) * rad["or2"], rtol=1e-12)   # orho = 1


def _compare(got, ref, mask, label):
    worst = 0.0
    for s in range(NDIAG):
        scale = np.abs(ref[:, s]).max()
        if scale == 0.0:
            assert np.abs(got[:, s]).max() == 0.0, f"{label}: slot {s} should be zero"
            continue
        err = np.abs(got[:, s] - ref[:, s]).max() / scale
        worst = max(worst, err)
        assert err < TOL, f"{label}: slot {s} deviates by {err:.2e}"
    print(f"{label}: worst slot error {worst:.2e}")


CASES = [
    # physics, l_max, n_r, ktopv, kbotv, anel, extra mask, omega_ic
    ("mhd", 21, 7, 2, 2, False, 0, 0.0),
    ("mhd", 16, 6, 1, 1, False, 0, 0.0),
    ("anel", 32, 6, 1, 2, True, 0, 0.0),
    ("mhd", 16, 5, 2, 2, False, 0, 3.5),
    ("mhd", 16, 5, 2, 1, False, DIAG_RMSBULK, 0.0),
]


@pytest.mark.gpu
@pytest.mark.parametrize("physics,l_max,n_r,ktopv,kbotv,anel,extra,omega_ic", CASES)
def test_gpu_diagnostics_against_the_oracle(physics, l_max, n_r, ktopv, kbotv, anel, extra, omega_ic):
    from magic_b200 import RadialLoop, Sht
    s = Sht(l_max)
    o = _oracle(l_max)
    p, rad, f = _case(physics, l_max, n_r, s.lm2l, s.lm2m, 11 + l_max, ktopv, kbotv, anel)
    if omega_ic:
        p.omega_ic, p.l_rot_ic = omega_ic, 1
    rl = RadialLoop(s, p, rad)
    for mask in (ALL | extra, DIAG_HEL | extra, DIAG_HEMI | DIAG_POWER | extra):
        got = rl.diagnostics(f, mask, ktops=1, kbots=2)
        ref = o.radial_diagnostics(_oparams(p), rad, f, mask, ktops=1, kbots=2)
        _compare(got, ref, mask, f"{physics} l{l_max} ktopv{ktopv} kbotv{kbotv} mask {mask}")
    again = rl.diagnostics(f, ALL | extra, ktops=1, kbots=2)
    assert np.array_equal(again, rl.diagnostics(f, ALL | extra, ktops=1, kbots=2)), "diagnostics are not bitwise repeatable"
    rl.finalize()
    s.finalize_sht()


@pytest.mark.gpu
def test_gpu_diagnostics_device_pointers_and_level_chunks():
    """Device-resident inputs give the bits of the host-pointer call; a radial loop run before and after is unaffected."""
    import torch
    from magic_b200 import RadialLoop, Sht
    l_max, n_r = 32, 40          # more levels than one diagnostics chunk (32): exercises the shifted last chunk
    s = Sht(l_max)
    p, rad, f = _case("mhd", l_max, n_r, s.lm2l, s.lm2m, 5)
    rl = RadialLoop(s, p, rad)
    before = rl.radialLoop(f)
    host = rl.diagnostics(f, ALL)
    dev = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in f.items()}
    got = rl.diagnostics({k: v.data_ptr() for k, v in dev.items()}, ALL, device=True)
    assert np.array_equal(host, got)
    after = rl.radialLoop(f)
    for k in before:
        assert np.array_equal(before[k], after[k]), k
    o = _oracle(l_max)
    _compare(host, o.radial_diagnostics(_oparams(p), rad, f, ALL), ALL, "mhd l32, 40 levels")
    rl.finalize()
    s.finalize_sht()


def test_oracle_dtB_products_against_the_per_call_transforms():
    """orc_radial_dtB (get_dtBLM, dtB.f90:144-223) on a bulk level against the same products formed in numpy from the oracle's
    own syntheses and analysed with its per-call transforms."""
    l_max, n_r = 16, 3     # n_theta = 24: the reference's grids have n_theta % 4 == 0 (truncation.f90:137-152)
    o = _oracle(l_max)
    p, rad, f = _case("mhd", l_max, n_r, o.lm2l, o.lm2m, 21, anel=True)
    out = o.radial_dtB(_oparams(p), rad, f)
    i = 1
    vr, vt, vp = o.torpol_to_spat(f["w"][i], f["dw"][i], f["z"][i], l_max)
    br, bt, bp = o.torpol_to_spat(f["b"][i], f["db"][i], f["aj"][i], l_max)
    n_t = len(o.theta_ord)             # N/S-interleaved colatitudes of the grid rows (row 2k = k-th northern node, 2k+1 its mirror)
    theta = np.empty(n_t)
    theta[0::2], theta[1::2] = o.theta_ord[:n_t // 2], o.theta_ord[::-1][:n_t // 2]
    os2, cot = 1.0 / np.sin(theta) ** 2, np.cos(theta) / np.sin(theta) ** 3
    rho = rad["orho1"][i]
    vpAS = rho * vp.mean(axis=0)
    s1, t1 = o.spat_to_sphertor(rho * bt * vr, rho * bp * vr, l_max)
    s2, t2 = o.spat_to_sphertor(rho * vt * br, rho * vp * br, l_max)
    want = [s1, t1, s2, t2, o.scal_to_SH(os2 * rho * bt * vp, l_max), o.scal_to_SH(os2 * rho * bp * vt, l_max),
            o.scal_to_SH(cot * rho * (bp * vt + bt * vp), l_max), o.scal_to_SH(os2 ** 2 * rho * (bp * vt + bt * vp), l_max),
            o.scal_to_SH(os2 * br * vpAS, l_max), o.scal_to_SH(os2 * bt * vpAS, l_max), o.scal_to_SH(os2 ** 2 * bt * vpAS, l_max)]
    for q in range(11):
        assert np.linalg.norm(out[q, i] - want[q]) < 1e-13 * np.linalg.norm(want[q]), q


@pytest.mark.gpu
@pytest.mark.parametrize("l_max,n_r,ktopv,kbotv,anel,omega_ic", [(21, 6, 2, 2, False, 0.0), (16, 5, 1, 2, True, 2.5), (32, 40, 2, 1, False, 0.0)])
def test_gpu_dtB_batch_against_the_oracle(l_max, n_r, ktopv, kbotv, anel, omega_ic):
    """magic_rloop_dtb (get_dtBLM as one more batch on the Legendre / FFT kernels) against the oracle: rigid and
    stress-free walls, a rotating inner core, an anelastic background, more levels than one chunk."""
    from magic_b200 import RadialLoop, Sht
    s = Sht(l_max)
    o = _oracle(l_max)
    p, rad, f = _case("mhd", l_max, n_r, s.lm2l, s.lm2m, 31 + l_max, ktopv, kbotv, anel)
    if omega_ic:
        p.omega_ic, p.l_rot_ic = omega_ic, 1
    rl = RadialLoop(s, p, rad)
    got = rl.dtb(f)
    ref = o.radial_dtB(_oparams(p), rad, f)
    names = ["BtVrLM", "BpVrLM", "BrVtLM", "BrVpLM", "BtVpLM", "BpVtLM", "BpVtBtVpCotLM", "BpVtBtVpSn2LM", "BrVZLM", "BtVZLM", "BtVZsn2LM"]
    for q, nm in enumerate(names):
        err = np.linalg.norm(got[q] - ref[q]) / np.linalg.norm(ref[q])
        print(f"  dtB l{l_max} {nm}: rel_l2 {err:.2e}")
        assert err < 1e-12, (nm, err)
    assert np.array_equal(got, rl.dtb(f)), "dtB batch is not bitwise repeatable"
    rl.finalize()
    s.finalize_sht()


@pytest.mark.gpu
def test_gpu_graph_fields_are_the_per_call_transforms():
    from magic_b200 import RadialLoop, Sht
    l_max, n_r = 16, 4
    s = Sht(l_max)
    o = _oracle(l_max)
    p, rad, f = _case("mhd", l_max, n_r, s.lm2l, s.lm2m, 9)
    rl = RadialLoop(s, p, rad)
    g = rl.graph_fields(f, 2, mag=True, pressure=True)
    vr, vt, vp = o.torpol_to_spat(f["w"][2], f["dw"][2], f["z"][2], l_max)
    br, bt, bp = o.torpol_to_spat(f["b"][2], f["db"][2], f["aj"][2], l_max)
    for got, ref in ((g["vr"], vr), (g["vt"], vt), (g["vp"], vp), (g["br"], br), (g["bt"], bt), (g["bp"], bp),
                     (g["sr"], o.scal_to_spat(f["s"][2], l_max)), (g["pr"], o.scal_to_spat(f["p"][2], l_max))):
        assert np.linalg.norm(got - ref) < 1e-12 * np.linalg.norm(ref)
    rl.finalize()
    s.finalize_sht()


@pytest.mark.gpu
def test_gpu_log_step_batches_share_one_workspace():
    """The diagnostics, get_dtBLM, TO and RMS batches are carved from one device workspace (they never run concurrently): used in
    turn on one plan, with the workspace growing and changing hands in between, each returns bit for bit what it returns on a
    plan of its own."""
    from magic_b200 import RadialLoop, Sht
    l_max, n_r = 21, 9
    s = Sht(l_max)
    p, rad, f = _case("mhd", l_max, n_r, s.lm2l, s.lm2m, 17)
    old = {k: 0.9 * f[k] for k in ("w", "dw", "z")}
    calls = {
        "hemi": lambda rl: rl.diagnostics(f, DIAG_HEMI),
        "all": lambda rl: rl.diagnostics(f, ALL),
        "dtb": lambda rl: rl.dtb(f),
        "to": lambda rl: (rl.to_next(f), rl.to(f, 1e-3))[1],
        "rms": lambda rl: (rl.rms_keep(old), rl.rms(f, 1e-3))[1],
    }
    alone = {}
    for nm, fn in calls.items():
        rl = RadialLoop(s, p, rad)
        alone[nm] = fn(rl)
        rl.finalize()
    rl = RadialLoop(s, p, rad)
    for nm in ("hemi", "dtb", "all", "rms", "to", "hemi", "rms", "dtb", "all", "to"):   # small -> large -> small requests, every hand-over
        assert np.array_equal(calls[nm](rl), alone[nm]), nm
    rl.finalize()
    s.finalize_sht()


def test_oracle_convective_flux_sums_obey_parseval():
    """get_fluxes (outPar.f90:470-582): the two grid sums of the convective flux, int vr s dOmega and int vr p dOmega (slots 23, 24;
    vr = r^2 u_r = sum l(l+1) w_lm Y_lm), equal sum_lm (2 - delta_m0) l(l+1) Re(w_lm conj(s_lm)) for an orthonormal basis --
    a closed form that does not go through the grid.  Bulk levels (on rigid walls vr is set to zero)."""
    l_max, n_r = 16, 5
    o = _oracle(l_max)
    p, rad, f = _case("mhd", l_max, n_r, o.lm2l, o.lm2m, 12)
    d = o.radial_diagnostics(_oparams(p), rad, f, DIAG_FLUX)
    fac = np.where(o.lm2m == 0, 1.0, 2.0)
    dLh = o.lm2l * (o.lm2l + 1.0)
    for i in range(1, n_r - 1):
        for slot, nm in ((23, "s"), (24, "p")):
            ref = np.sum(fac * dLh * (f["w"][i] * np.conj(f[nm][i])).real)
            assert abs(d[i, slot] - ref) < 1e-12 * max(abs(ref), np.abs(f["w"][i]).max() * np.abs(f[nm][i]).max() * dLh.max())
    assert not d[0, 23] and not d[-1, 23]


def test_oracle_boundary_layer_and_perpPar_sums_against_spectral_forms():
    """get_nlBLayers (outPar.f90:584-644): gradT2ASr = (1 / 2 pi) int |grad s|^2 dOmega = (1 / 2 pi) sum (2 - delta_m0) (|ds_lm|^2 +
    l(l+1) |s_lm|^2 / r^2); get_perpPar (outPar.f90:646-726): E_perp + E_par is the kinetic energy density that get_hemi sums
    (golden-pinned through hemi.TAG), 2 pi (EperpASr + EparASr) = (ekin_N + ekin_S) / r^2 for orho = 1."""
    l_max, n_r = 16, 5
    o = _oracle(l_max)
    p, rad, f = _case("mhd", l_max, n_r, o.lm2l, o.lm2m, 21)
    d = o.radial_diagnostics(_oparams(p), rad, f, DIAG_VISCBC | DIAG_PERPPAR | DIAG_HEMI | DIAG_RMSBULK, ktops=2, kbots=2)
    fac = np.where(o.lm2m == 0, 1.0, 2.0)
    dLh = o.lm2l * (o.lm2l + 1.0)
    for i in range(n_r):
        ref = (np.sum(fac * np.abs(f["ds"][i]) ** 2) + rad["or2"][i] * np.sum(fac * dLh * np.abs(f["s"][i]) ** 2)) / (2.0 * np.pi)
        assert abs(d[i, 30] - ref) < 1e-12 * ref
        np.testing.assert_allclose(2.0 * np.pi * (d[i, 18] + d[i, 19]), (d[i, 9] + d[i, 10]) * rad["or2"][i], rtol=1e-12)
        assert d[i, 20] <= d[i, 18] and d[i, 21] <= d[i, 19] + 1e-15          # the axisymmetric parts are parts
