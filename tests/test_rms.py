"""R.m.s. force balance, the part inside the radial loop on lRmsCalc steps (rIter.f90:215-252, 710; RMS.f90:469-610;
SURVEY.md 8(f)4): transform_to_grid_RMS, get_nl with every level as bulk, get_nl_RMS, transform_to_lm_RMS.

CPU: the oracle's restatement (oracle/magic_oracle_diag.inc orc_radial_RMS) against
  * closed forms in spectral space: with the kept velocity a scaled copy of the present one, dtVrLM = f or2 l(l+1) w and
    (dtVtLM, dtVpLM) = f or1 (dw, z); without the curl-form correction (PFt2LM, PFp2LM) = (or1 p, 0),
  * the same grid products formed in numpy from the oracle's golden-pinned per-call transforms and analysed with the per-call
    analyses (Coriolis, pressure-gradient and advection terms in both forms, dpkindr, the Lorentz terms, the merged AdvrLM).
GPU (-m gpu): magic_rloop_rms_keep / magic_rloop_rms through the C ABI against the oracle on the same seeded spectra (MHD in curl
form, anelastic hydro in u.grad u form, Boussinesq hydro, MHD with precession + centrifugal acceleration + a phase field; more
levels than one chunk; device pointers).
"parity unpinned": samples/testRMSOutputs compares dtVrms.TAG / dtBrms.TAG, which need compute_lm_forces and the radial
integration of dtVrms on the host.
"""
import numpy as np
import pytest

from magic_b200.riter import NRMS
from magic_b200.workload import make_fields, make_params, make_radial

NAMES = ["AdvrLM", "LFrLM", "dtVrLM", "dpkindrLM", "Advt2LM", "Advp2LM", "LFt2LM", "LFp2LM", "CFt2LM", "CFp2LM", "PFt2LM", "PFp2LM", "dtVtLM",
         "dtVpLM"]


def _oracle(l_max, minc=1):
    from oracle.oracle import Oracle
    return Oracle(l_max, minc=minc)


def _oparams(p):
    from oracle.oracle import Params as OParams
    op = OParams()
    for n, _ in p._fields_:
        setattr(op, n, getattr(p, n))
    return op


def _case(physics, l_max, n_r_max, lm2l, lm2m, seed, anel=False, extra=False, **kw):
    p = make_params(physics, n_r_max, **kw)
    rad = make_radial(n_r_max, l_max, anel=anel)
    f = make_fields(physics, lm2l, lm2m, n_r_max, seed)
    if extra:   # precession, centrifugal acceleration and a phase field: the terms that only enter AdvrLM / the penalty
        p.l_precession, p.po, p.prec_angle, p.oek = 1, 0.3, 0.4, 1e3
        p.l_centrifuge, p.dilution_fac, p.ra, p.opr = 1, 0.02, 1e5, 1.0
        p.l_phase_field, p.epsPhase, p.penaltyFac, p.phaseDiffFac, p.tmelt = 1, 0.05, 0.7, 1.0, 0.3
        f["phi"] = 0.4 * f["s"] + 0.2 * f["w"]
    rng = np.random.default_rng(seed + 3)
    f["p"] = f["s"] * (0.3 + rng.random()) + 0.1 * f["w"]
    old = {k: 0.7 * f[k] + 0.2 * f[q] for k, q in (("w", "z"), ("dw", "dz"), ("z", "w"))}
    return p, rad, f, old


def test_oracle_rms_closed_forms_in_spectral_space():
    l_max, n_r = 16, 4
    o = _oracle(l_max)
    p, rad, f, _ = _case("anel", l_max, n_r, o.lm2l, o.lm2m, 2)     # u.grad u form: no curl-form correction of the pressure term
    dt, fac = 2e-4, 0.25
    old = {k: (1.0 - fac) * f[k] for k in ("w", "dw", "z")}
    got = o.radial_RMS(_oparams(p), rad, f, old, dt)
    dLh = (o.lm2l * (o.lm2l + 1.0))[None, :]
    or1, or2 = rad["or1"][:, None], rad["or2"][:, None]
    ref = {"dtVrLM": fac / dt * or2 * dLh * f["w"], "dtVtLM": fac / dt * or1 * f["dw"], "dtVpLM": fac / dt * or1 * f["z"],
           "PFt2LM": or1 * f["p"], "PFp2LM": 0.0 * f["p"], "dpkindrLM": 0.0 * f["p"], "LFrLM": 0.0 * f["p"]}
    for nm, r in ref.items():
        r = np.where(o.lm2l[None, :] == 0, 0.0, r) if nm != "dtVrLM" else r
        g = got[NAMES.index(nm)]
        assert np.abs(g - r).max() <= 1e-12 * max(np.abs(r).max(), np.abs(f["p"]).max()), nm


@pytest.mark.parametrize("physics,anel,extra", [("mhd", False, False), ("anel", True, False), ("hydro", False, True)])
def test_oracle_rms_against_products_of_the_per_call_transforms(physics, anel, extra):
    l_max, n_r = 16, 4
    o = _oracle(l_max)
    p, rad, f, old = _case(physics, l_max, n_r, o.lm2l, o.lm2m, 9, anel=anel, extra=extra)
    p.CorFac = 321.0
    dt, time = 1e-3, 0.37
    got = o.radial_RMS(_oparams(p), rad, f, old, dt, time=time)
    st, ct = np.sin(np.arccos(o.cosTheta))[None, :], o.cosTheta[None, :]          # grids are [n_phi, n_theta]
    os2, cn2 = 1.0 / st ** 2, ct / st ** 2
    for i in (0, 2):                                                               # a boundary level (bulk with lRmsCalc) and a bulk one
        r, or1, or2, orho1, beta = (rad[k][i] for k in ("r", "or1", "or2", "orho1", "beta"))
        or3, or4 = or1 * or2, or2 * or2
        vr, vt, vp = o.torpol_to_spat(f["w"][i], f["dw"][i], f["z"][i], l_max)
        dvrdr, dvtdr, dvpdr = o.torpol_to_spat(f["dw"][i], f["ddw"][i], f["dz"][i], l_max)
        cvr, cvt, cvp = o.torpol_to_curl_spat(or2, f["w"][i], f["ddw"][i], f["z"][i], f["dz"][i], l_max)
        dvrdt, dvrdp = o.pol_to_grad_spat(f["w"][i], l_max)
        dvtdp, dvpdp = o.torpol_to_dphspat(f["dw"][i], f["z"][i], l_max)
        dpdt, dpdp = o.scal_to_grad_spat(f["p"][i], l_max)
        vro, vto, vpo = o.torpol_to_spat(old["w"][i], old["dw"][i], old["z"][i], l_max)
        if p.l_adv_curl:
            Ar, At, Ap = -os2 * (cvt * vp - cvp * vt), -or4 * (cvp * vr - cvr * vp), -or4 * (cvr * vt - cvt * vr)
        else:
            Ar = -or2 * orho1 * (vr * (dvrdr - (2 * or1 + beta) * vr) + os2 * (vt * (dvrdt - r * vt) + vp * (dvrdp - r * vp)))
            At = or4 * orho1 * (-vr * (dvtdr - beta * vt) + vt * (cn2 * vt + dvpdp + dvrdr) + vp * (cn2 * vp - dvtdp))
            Ap = or4 * orho1 * (-vr * (dvpdr - beta * vp) - vt * (dvtdp + cvr) - vp * dvpdp)
        if p.l_phase_field:                                                        # get_nl.f90:333-339
            phi = o.scal_to_spat(f["phi"][i], l_max)
            pen = 1.0 / p.epsPhase ** 2 / p.penaltyFac ** 2
            Ar, At, Ap = Ar - phi * vr * pen, At - or2 * phi * vt * pen, Ap - or2 * phi * vp * pen
        PFt, PFp, At2, Ap2 = or1 * dpdt, or1 * dpdp, r * At, r * Ap
        ref = {}
        if p.l_adv_curl:
            X = or3 * (or2 * vr * dvrdt - vt * (dvrdr + dvpdp + cn2 * vt) + vp * (cvr + dvtdp - cn2 * vp))
            Y = or3 * (or2 * vr * dvrdp + vt * dvtdp + vp * dvpdp)
            PFt, PFp, At2, Ap2 = PFt - X, PFp - Y, At2 - X, Ap2 - Y
            ref["dpkindrLM"] = o.scal_to_SH(or4 * vr * (dvrdr - 2 * or1 * vr) + or2 * os2 * (vt * (dvtdr - or1 * vt) + vp * (dvpdr - or1 * vp)), l_max)
        ref["PFt2LM"], ref["PFp2LM"] = o.spat_to_sphertor(PFt, PFp, l_max)
        ref["Advt2LM"], ref["Advp2LM"] = o.spat_to_sphertor(At2, Ap2, l_max)
        ref["CFt2LM"], ref["CFp2LM"] = o.spat_to_sphertor(-2 * p.CorFac * ct * vp * or1, 2 * p.CorFac * st * (or1 * ct / st * vt + or2 * st * vr), l_max)
        ref["dtVrLM"] = o.scal_to_SH(or2 * (vr - vro) / dt, l_max)
        ref["dtVtLM"], ref["dtVpLM"] = o.spat_to_sphertor(or1 * (vt - vto) / dt, or1 * (vp - vpo) / dt, l_max)
        if p.l_mag_LF:
            br, bt, bp = o.torpol_to_spat(f["b"][i], f["db"][i], f["aj"][i], l_max)
            cbr, cbt, cbp = o.torpol_to_curl_spat(or2, f["b"][i], f["ddb"][i], f["aj"][i], f["dj"][i], l_max)
            LFr, LFt, LFp = p.LFfac * os2 * (cbt * bp - cbp * bt), p.LFfac * or4 * (cbp * br - cbr * bp), p.LFfac * or4 * (cbr * bt - cbt * br)
            ref["LFrLM"] = o.scal_to_SH(LFr, l_max)
            ref["LFt2LM"], ref["LFp2LM"] = o.spat_to_sphertor(r * LFt, r * LFp, l_max)
            Ar = Ar + LFr
        if p.l_precession:                                                         # get_nl.f90:346-357, merged at rIter.f90:669-673
            ph = (p.oek * time + 2.0 * np.pi * np.arange(o.n_phi) / (o.n_phi * o.minc))[:, None]
            Ar = Ar - 2.0 * p.oek * p.po * np.sin(p.prec_angle) / st * r * (np.cos(ph) * vp * ct + np.sin(ph) * vt)
        if p.l_centrifuge:                                                         # get_nl.f90:359-367, merged at rIter.f90:675-678
            Ar = Ar - p.dilution_fac * r * st ** 4 * p.ra * p.opr * o.scal_to_spat(f["s"][i], l_max)
        ref["AdvrLM"] = o.scal_to_SH(Ar, l_max)
        for nm, rv in ref.items():
            g = got[NAMES.index(nm), i]
            partner = nm.replace("t2LM", "p2LM") if "t2LM" in nm else nm.replace("p2LM", "t2LM")     # scale of a (spheroidal, toroidal) pair
            scale = max(np.abs(rv).max(), np.abs(ref.get(partner, rv)).max())
            assert np.abs(g - rv).max() < 1e-12 * scale, (nm, i)
        for nm in NAMES:
            if nm not in ref:
                assert not got[NAMES.index(nm), i].any(), nm


def _compare(got, ref, label, tol=1e-12):
    worst = 0.0
    for q in range(NRMS):
        q2 = q if q < 4 else (q + 1 if q % 2 == 0 else q - 1)     # the partner of a (spheroidal, toroidal) pair sets the scale
        scale = max(np.abs(ref[q]).max(), np.abs(ref[q2]).max())
        if np.abs(ref[q]).max() == 0.0:
            assert np.abs(got[q]).max() == 0.0, f"{label}: {NAMES[q]} should be zero"
            continue
        err = np.abs(got[q] - ref[q]).max() / scale
        worst = max(worst, err)
        assert err < tol, f"{label}: {NAMES[q]} deviates by {err:.2e}"
    print(f"{label}: worst array error {worst:.2e}")


@pytest.mark.gpu
@pytest.mark.parametrize("physics,l_max,n_r,anel,ktopv,extra", [("mhd", 21, 7, False, 2, False), ("anel", 16, 6, True, 1, False),
                                                                 ("hydro", 32, 40, False, 2, False), ("mhd", 16, 6, False, 2, True)])
def test_gpu_rms_batch_matches_oracle(physics, l_max, n_r, anel, ktopv, extra):
    from magic_b200 import RadialLoop, Sht
    s = Sht(l_max)
    o = _oracle(l_max)
    p, rad, f, old = _case(physics, l_max, n_r, s.lm2l, s.lm2m, 31, anel=anel, extra=extra, ktopv=ktopv, kbotv=ktopv)
    dt, time = 4e-4, 0.37
    ref = o.radial_RMS(_oparams(p), rad, f, old, dt, time=time)
    rl = RadialLoop(s, p, rad)
    rl.radialLoop(f, time=time)            # the batch follows a pass of the loop and takes its time for the precession terms
    with pytest.raises(Exception):
        rl.rms(f, dt)                      # nothing kept yet: loud
    rl.rms_keep(old)
    got = rl.rms(f, dt)
    _compare(got, ref, f"RMS {physics} l{l_max}")
    assert np.array_equal(got, rl.rms(f, dt))
    rl.rms_keep(f)                         # the next step: nothing has moved, the time derivative vanishes up to round-off
    again = rl.rms(f, dt)
    assert np.abs(again[2]).max() < 1e-9 * np.abs(got[2]).max() and np.array_equal(again[4], got[4])
    rl.finalize()
    s.finalize_sht()


@pytest.mark.gpu
def test_gpu_rms_with_device_pointers():
    import torch
    from magic_b200 import RadialLoop, Sht
    l_max, n_r = 16, 6
    s = Sht(l_max)
    p, rad, f, old = _case("mhd", l_max, n_r, s.lm2l, s.lm2m, 8)
    rl = RadialLoop(s, p, rad)
    rl.rms_keep(old)
    ref = rl.rms(f, 1e-3)
    dev = lambda d: {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in d.items()}
    kf, ko = dev(f), dev(old)
    torch.cuda.synchronize()
    rl.rms_keep({k: v.data_ptr() for k, v in ko.items()}, device=True)
    got = rl.rms({k: v.data_ptr() for k, v in kf.items()}, 1e-3, device=True)
    assert np.array_equal(got, ref)
    rl.finalize()
    s.finalize_sht()
