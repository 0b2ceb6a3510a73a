"""Synthetic workloads of the radial-loop benchmark (SURVEY.md 8d / BASELINE.md section 2).

Host-side helper shared by bench.py and the parity tests: builds the run-wide switches, the radial
functions of a Chebyshev shell (radratio 0.35) and seeded random spectra for one rank's radial slab.
"""
import numpy as np

from .riter import Params
from .sht import grid_sizes
from .transpose import get_blocks

CONFIGS = {
    # name: (l_max, minc, n_r_max, n_phi_tot or None, physics)
    "dynamo_benchmark": dict(l_max=16, minc=1, n_r_max=33, physics="mhd", config_id=0),
    "hydro_bench_anel": dict(l_max=0, n_phi_tot=288, minc=1, n_r_max=97, physics="anel", config_id=1),
    # BASELINE.json quotes this case at "l_max=85": the grid when l_max is imposed instead of n_phi_tot
    "hydro_bench_anel_l85": dict(l_max=85, minc=1, n_r_max=97, physics="anel", config_id=1),
    "bouss_dynamo_l255": dict(l_max=255, minc=1, n_r_max=121, physics="mhd", config_id=2),
    # samples/full_sphere scaled up: FD radial scheme => double-curl poloidal equation, r = 0 level (v_center_sphere),
    # l_R(r) falling towards the centre (l_var_l, radial.f90:291-307 with rcut_l = 0.1)
    "full_sphere_l511": dict(l_max=511, minc=1, n_r_max=161, physics="hydro", config_id=3,
                             flags=dict(l_full_sphere=1, l_double_curl=1), l_var_l=True),
    "dynamo_l1023": dict(l_max=1023, minc=1, n_r_max=257, physics="mhd", config_id=4),
    # the radial loop on a real saturated state -- the spectra of samples/boussBenchSat/checkpoint_end.start (read
    # with magic_b200.checkpoint and committed as tests/golden/boussBenchSat_ckpt.npz), radial derivatives by the Chebyshev
    # collocation matrices of the same grid; l_max = 64, minc = 4, n_r_max = 33
    "boussBenchSat_ckpt": dict(l_max=64, minc=4, n_r_max=33, physics="mhd", config_id=5,
                               checkpoint="tests/golden/boussBenchSat_ckpt.npz"),
}


def make_params(physics, n_r_max, ktopv=2, kbotv=2):
    """Switches of the reference cases: 'mhd' = Boussinesq MHD with curl-form advection (dynamo_benchmark),
    'hydro' = Boussinesq hydro + heat, 'anel' = anelastic hydro with u.grad u advection (hydro_bench_anel)."""
    p = Params()
    p.l_conv = 1
    p.l_heat = 1
    p.l_conv_nl = 1
    p.l_heat_nl = 1
    p.l_corr = 1
    p.l_adv_curl = 1
    p.l_cour_alf_damp = 1
    p.ktopv, p.kbotv = ktopv, kbotv
    p.n_r_max = n_r_max
    p.n_r_LCR = 0
    ek, pm = 1e-3, 5.0
    p.CorFac = 1.0 / ek
    p.epsc = 0.0
    p.opm = 1.0 / pm
    p.courfac, p.alffac = 2.5, 1.0
    p.r_cmb, p.r_icb = 20.0 / 13.0, 7.0 / 13.0
    p.LFfac = 1.0 / (ek * pm)
    if physics == "mhd":
        p.l_mag = p.l_mag_nl = p.l_mag_LF = 1
    elif physics == "anel":
        p.l_anel = 1
        p.l_adv_curl = 0
        p.ViscHeatFac = 0.01
        p.OhmLossFac = 0.0
    elif physics != "hydro":
        raise ValueError(physics)
    return p


def make_radial(n_r_max, l_max, nRstart=1, nRstop=None, l_R=None, anel=False):
    """Chebyshev (Gauss-Lobatto) radii, nR=1 is the CMB; slice [nRstart, nRstop] (1-based inclusive)."""
    if nRstop is None:
        nRstop = n_r_max
    r_cmb, r_icb = 20.0 / 13.0, 7.0 / 13.0
    j = np.arange(n_r_max)
    r = r_icb + (r_cmb - r_icb) * 0.5 * (1.0 + np.cos(np.pi * j / (n_r_max - 1)))
    lR = np.full(n_r_max, l_max, dtype=np.int32) if l_R is None else np.asarray(l_R, dtype=np.int32)
    delxr2 = np.zeros(n_r_max)  # preCalculations.f90:304-310
    delxr2[0] = (r[0] - r[1]) ** 2
    delxr2[-1] = (r[-2] - r[-1]) ** 2
    for n in range(1, n_r_max - 1):
        delxr2[n] = min(r[n - 1] - r[n], r[n] - r[n + 1]) ** 2
    delxh2 = r ** 2 / (lR * (lR + 1.0))
    one = np.ones(n_r_max)
    if anel:  # a smooth polytropic-like background: rho0 = (1 + 0.5 (r_cmb - r))^2
        rho0 = (1.0 + 0.5 * (r_cmb - r)) ** 2
        beta = -1.0 / (1.0 + 0.5 * (r_cmb - r))  # dln(rho0)/dr
        temp0 = 1.0 + 0.5 * (r_cmb - r)
    else:
        rho0, beta, temp0 = one, 0 * one, one
    full = dict(nR=np.arange(1, n_r_max + 1, dtype=np.int32), l_R=lR, r=r, or1=1 / r, or2=1 / r ** 2, or4=1 / r ** 4,
                orho1=1 / rho0, orho2=1 / rho0 ** 2, beta=beta, rho0=rho0, otemp1=1 / temp0, temp0=temp0, visc=one,
                epscProf=one, delxr2=delxr2, delxh2=delxh2)
    full["lambda"] = one
    sl = slice(nRstart - 1, nRstop)
    return {k: np.ascontiguousarray(v[sl]) for k, v in full.items()}


def cheb_matrices(n_r_max, r_icb=7.0 / 13.0, r_cmb=20.0 / 13.0):
    """First and second derivative collocation matrices on the Gauss-Lobatto radii of make_radial (nR = 1 is the CMB): what
    get_dr / get_ddr (radial_derivatives.f90:714-912) compute on this grid, as the dense matrices
    magic_rloop_set_radial_matrices takes."""
    N = n_r_max - 1
    x = np.cos(np.pi * np.arange(n_r_max) / N)
    c = np.ones(n_r_max)
    c[0] = c[-1] = 2.0
    c *= (-1.0) ** np.arange(n_r_max)
    X = np.tile(x, (n_r_max, 1)).T
    dX = X - X.T
    D = np.outer(c, 1.0 / c) / (dX + np.eye(n_r_max))
    D -= np.diag(D.sum(axis=1))
    D *= 2.0 / (r_cmb - r_icb)
    return D, D @ D


FIELD_SETS = {
    "mhd": ["w", "dw", "ddw", "z", "dz", "s", "b", "db", "ddb", "aj", "dj"],
    "hydro": ["w", "dw", "ddw", "z", "dz", "s"],
    "anel": ["w", "dw", "ddw", "z", "dz", "s"],
}


def make_fields(physics, lm2l, lm2m, n_r_loc, seed, out=None):
    """Re,Im ~ N(0,1)/(l+1), Im=0 at m=0, l=0 entries of all fields zero except s."""
    rng = np.random.default_rng(seed)
    lm_max = len(lm2l)
    scale = 1.0 / (lm2l + 1.0)
    fields = {} if out is None else out
    for nm in FIELD_SETS[physics]:
        a = fields.get(nm)
        if a is None:
            a = np.empty((n_r_loc, lm_max), dtype=np.complex128)
            fields[nm] = a
        for i in range(n_r_loc):
            re = rng.standard_normal(lm_max)
            im = rng.standard_normal(lm_max)
            im[lm2m == 0] = 0.0
            a[i] = (re + 1j * im) * scale
        if nm != "s":
            a[:, lm2l == 0] = 0.0
    return fields


def config_sizes(name):
    c = CONFIGS[name]
    gs = grid_sizes(l_max=c["l_max"], n_phi_tot=c.get("n_phi_tot", 0), minc=c["minc"])
    gs.update(n_r_max=c["n_r_max"], minc=c["minc"], physics=c["physics"], config_id=c["config_id"], flags=c.get("flags", {}),
              l_var_l=c.get("l_var_l", False), checkpoint=c.get("checkpoint"))
    return gs


def config_params(gs):
    """make_params + the configuration's extra switches."""
    p = make_params(gs["physics"], gs["n_r_max"])
    for k, v in gs.get("flags", {}).items():
        setattr(p, k, v)
    return p


def config_l_R(gs):
    """l_R(nR) = min(l_max, int(1 + l_max sqrt(x / rcut_l))) of radial.f90:291-296 (l_var_l, rcut_l = 0.1), with x the distance
    from the innermost level as a fraction of the shell depth (the full sphere's r / r_cmb), or None for l_R = l_max."""
    if not gs.get("l_var_l"):
        return None
    n_r_max, l_max = gs["n_r_max"], gs["l_max"]
    r = make_radial(n_r_max, l_max)["r"]
    r_cmb = r[0]
    l_R = np.full(n_r_max, l_max, dtype=np.int32)
    for n in range(n_r_max):
        l_R[n] = min(l_max, int(1.0 + l_max * np.sqrt((r[n] - r[-1]) / (r_cmb - r[-1]) / 0.1)))
    return l_R


def seed_for(config_id, rank):
    return 20261017 + 1000 * config_id + rank


def checkpoint_containers(path, lm_max, n_r_max):
    """R-distributed containers (all levels) of a checkpoint fixture: flow = (w, dw, ddw, z, dz), s = (s, ds), field = (b, db, ddb,
    aj, dj), complex128 [nf, n_r_max, lm_max]; the radial derivatives are what get_dr / get_ddr return on the Chebyshev grid."""
    d = np.load(path)
    assert d["w"].shape == (n_r_max, lm_max), (d["w"].shape, n_r_max, lm_max)
    D1, D2 = cheb_matrices(n_r_max)
    dr = lambda D, a: np.einsum("ij,jk->ik", D, a)
    w, z, s, b, aj = (np.asarray(d[k], dtype=np.complex128) for k in ("w", "z", "s", "b", "aj"))
    return {"flow": np.stack([w, dr(D1, w), dr(D2, w), z, dr(D1, z)]), "s": np.stack([s, dr(D1, s)]),
            "field": np.stack([b, dr(D1, b), dr(D2, b), aj, dr(D1, aj)])}
