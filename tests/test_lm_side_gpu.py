"""GPU parity of the LM-side prologue / epilogue of magic_rloop_run_lm.

The host-container call can (a) compute the radial derivatives dw, ddw, dz, db, ddb, dj on the device from w, z, b, aj with the
host's radial matrices (get_dr / get_ddr, radial_derivatives.f90:714-912, as dense matrices) instead of receiving them over
PCIe, and (b) run finish_explicit_assembly (LMLoop.f90:390-453: finish_exp_entropy updateS.f90:543-601, finish_exp_mag
updateB.f90:1005-1041, finish_exp_pol updateWP.f90:1002-1031) on the device after the outbound transposes.  Both are checked
against a numpy restatement of those formulas applied to the results of the plain call; the matrices are the Chebyshev
collocation derivative matrices of oracle/lmloop.py's grid (any matrix would do: the library only multiplies).
One rank: the transposes are the lo <-> st permutation of every level chunk.
"""
import numpy as np
import pytest

from tests.util import rel_l2

pytestmark = pytest.mark.gpu


def _cheb_matrices(n_r, r_icb=7.0 / 13.0, r_cmb=20.0 / 13.0):
    """First and second derivative collocation matrices on the Gauss-Lobatto radii (nR = 1 is the CMB)."""
    N = n_r - 1
    x = np.cos(np.pi * np.arange(n_r) / N)
    c = np.ones(n_r)
    c[0] = c[-1] = 2.0
    c *= (-1.0) ** np.arange(n_r)
    X = np.tile(x, (n_r, 1)).T
    dX = X - X.T
    D = np.outer(c, 1.0 / c) / (dX + np.eye(n_r))
    D -= np.diag(D.sum(axis=1))
    D *= 2.0 / (r_cmb - r_icb)
    return D, D @ D


def _setup(physics, n_r=17, l_max=21, double_curl=False, level_chunk=4):
    from magic_b200 import RadialLoop, Sht, Transposer
    from magic_b200.transpose import lo_map
    from magic_b200.workload import make_fields, make_params, make_radial
    sht = Sht(l_max)
    tr = Transposer(sht, n_r, 5)
    p = make_params(physics, n_r)
    if double_curl:
        p.l_double_curl = 1
    l_R = np.full(n_r, l_max, dtype=np.int32)
    l_R[-5:] = [19, 16, 12, 7, 3]
    rad = make_radial(n_r, l_max, l_R=l_R)
    rl = RadialLoop(sht, p, rad, level_chunk=level_chunk)
    lo2st, _, _ = lo_map(l_max, l_max, 1, 1)
    g = make_fields(physics, sht.lm2l, sht.lm2m, n_r, 5)
    D1, D2 = _cheb_matrices(n_r)

    def lm(a):  # [n_r, lm_max] st order -> [n_r, nlm] lo order
        return np.ascontiguousarray(a[:, lo2st])
    w, z, s = lm(g["w"]), lm(g["z"]), lm(g["s"])
    dr = lambda D, a: np.einsum("ij,jk->ik", D, a)
    h_in = {"flow": np.stack([w, dr(D1, w), dr(D2, w), z, dr(D1, z)]), "s": np.stack([s, np.zeros_like(s)])}
    if physics == "mhd":
        b, aj = lm(g["b"]), lm(g["aj"])
        h_in["field"] = np.stack([b, dr(D1, b), dr(D2, b), aj, dr(D1, aj)])
    return sht, tr, p, rad, rl, h_in, D1, D2, sht.lm2l[lo2st], sht.lm2m[lo2st], l_R


def _run(rl, tr, h_in, p, n_r, nlm):
    out = {"dflowdt": np.zeros((4 if p.l_double_curl else 3, n_r, nlm), dtype=np.complex128), "dsdt": np.zeros((2, n_r, nlm), dtype=np.complex128)}
    if p.l_mag:
        out["dbdt"] = np.zeros((3, n_r, nlm), dtype=np.complex128)
    dtr, dth = np.zeros(n_r), np.zeros(n_r)
    rl.run_lm(tr, h_in, out, dtr, dth)
    return out, dtr, dth


@pytest.mark.parametrize("physics,double_curl", [("mhd", False), ("hydro", True)])
def test_derivatives_on_device(physics, double_curl):
    sht, tr, p, rad, rl, h_in, D1, D2, lo2l, lo2m, l_R = _setup(physics, double_curl=double_curl)
    n_r, nlm = h_in["flow"].shape[1:]
    ref, dtr_ref, _ = _run(rl, tr, h_in, p, n_r, nlm)
    rl.set_radial_matrices(D1, D2)
    rl.lm_options(derivs_on_device=True)
    poisoned = {k: v.copy() for k, v in h_in.items()}
    for k in ("flow", "field"):
        if k in poisoned:
            poisoned[k][[1, 2, 4]] = np.nan   # the derivative slots of the host containers must not be read
    got, dtr, _ = _run(rl, tr, poisoned, p, n_r, nlm)
    for k in ref:
        for f in range(ref[k].shape[0]):
            if np.linalg.norm(ref[k][f]) == 0:
                continue
            err = rel_l2(got[k][f], ref[k][f])
            print(f"  {physics} derivs on device: {k}[{f}] rel_l2 {err:.2e}")
            assert np.isfinite(got[k][f]).all() and err < 1e-11, (k, f, err)
    assert np.allclose(dtr, dtr_ref, rtol=1e-12)
    rl.finalize(); tr.destroy_comm(); sht.finalize_sht()


@pytest.mark.parametrize("physics,double_curl", [("mhd", False), ("hydro", True)])
def test_finish_explicit_assembly_on_device(physics, double_curl):
    sht, tr, p, rad, rl, h_in, D1, D2, lo2l, lo2m, l_R = _setup(physics, double_curl=double_curl)
    n_r, nlm = h_in["flow"].shape[1:]
    ref, _, _ = _run(rl, tr, h_in, p, n_r, nlm)
    rng = np.random.default_rng(3)
    or2, orho1 = rad["or2"], 1.0 + 0.1 * rng.random(n_r)
    dentropy0 = rng.standard_normal(n_r)
    # numpy restatement of finish_exp_entropy / finish_exp_mag / finish_exp_pol on the plain call's results
    dr = lambda a: np.einsum("ij,jk->ik", D1, a)
    on = lo2l[None, :] <= l_R[:, None]
    dL = (lo2l * (lo2l + 1.0))[None, :]
    want = {k: v.copy() for k, v in ref.items()}
    w = h_in["flow"][0]
    fin = (orho1[:, None] * (ref["dsdt"][0] - or2[:, None] * dr(ref["dsdt"][1]) - dL * (or2 * dentropy0)[:, None] * w))
    want["dsdt"][0] = np.where(on, fin, ref["dsdt"][0])
    if p.l_mag:
        fin = ref["dbdt"][1] + or2[:, None] * dr(ref["dbdt"][2])
        want["dbdt"][1] = np.where(on & ~((lo2l == 0) & (lo2m == 0))[None, :], fin, ref["dbdt"][1])
    if double_curl:
        fin = ref["dflowdt"][0] + or2[:, None] * dr(ref["dflowdt"][3])
        want["dflowdt"][0] = np.where(on & (lo2l > 0)[None, :], fin, ref["dflowdt"][0])
    rl.set_radial_matrices(D1, D2)
    rl.set_lm_radial(or2, orho1, dentropy0, l_R)
    rl.lm_options(finish_on_device=True)
    got, _, _ = _run(rl, tr, h_in, p, n_r, nlm)
    consumed = {("dsdt", 1), ("dbdt", 2), ("dflowdt", 3)}   # stay on the device: not written to the host arrays
    for k in ref:
        for f in range(ref[k].shape[0]):
            if (k, f) in consumed:
                assert not got[k][f].any(), (k, f)
                continue
            if np.linalg.norm(want[k][f]) == 0:
                continue
            err = rel_l2(got[k][f], want[k][f])
            print(f"  {physics} finish on device: {k}[{f}] rel_l2 {err:.2e}")
            assert err < 1e-12, (k, f, err)
    rl.finalize(); tr.destroy_comm(); sht.finalize_sht()
