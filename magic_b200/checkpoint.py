"""Reader / writer of MagIC checkpoint files (`checkpoint_*.TAG`), SURVEY.md 8(f)3.

Host-side mirror of `storeCheckPoints.f90:45-277` (`store`, format version 5) and of the sequential reader
`readCheckPoints.f90:767-1468` (`readStartFields`; the same layout is documented by `python/magic/checkpoint.py:165-335`):
a stream-access binary without record markers, little endian, holding the spectral state in `st_map` order -- the layout
the radial loop's R-distributed containers use (`complex128 [n_r_max][lm_max]`, lm fastest) -- so a checkpoint field can be
handed to `RadialLoop` / `Transposer` without any reordering.  No remapping between grids or truncations is done here
(`mapDataR`, `getLm2lmO` stay with the Fortran host); a checkpoint is read as it was written.

Layout (version 5; differences of older versions in brackets):
    int32    version
    float64  time
    char[10] time-scheme family ('MULTISTEP' or 'DIRK'), int32 nexp, nimp, nold
    float64  dt[nexp]  (MULTISTEP)  |  dt[1]  (DIRK)
    int32    n_time_step
    float64  ra, pr, raxi, sc, prmag, ek, stef, radratio, sigma_ratio            [version <= 2: no stef]
    int32    n_r_max, n_theta_max, n_phi_tot, minc, nalias, n_r_ic_max
    int32    l_max, m_min, m_max                                                 [version <= 3: absent, derived]
    char[72] radial scheme ('cheb' | 'fd'), int32 n_max|order, map|order_boundary, float64 ratio1, ratio2
    float64  r[n_r_max]
    MULTISTEP: float64 domega_ic_dt expl[2..nexp], impl[2..nimp], old[2..nold]; the same for domega_ma_dt
                                                                                 [version < 5: then lorentz_torque_ic/ma likewise]
    float64  omega_ic1, omegaOsz_ic1, tOmega_ic1, omega_ic2, omegaOsz_ic2, tOmega_ic2, and the same six for the mantle
    int32    l_heat, l_chemical_conv, l_phase_field, l_mag, l_press_store, l_cond_ic   [version <= 2: no l_phase_field]
    per field, in the order w, z, [p], [s], [xi], [phi], [b, aj], [b_ic, aj_ic]:
        complex128 field[n_r][lm_max], then (MULTISTEP) expl[2..nexp], impl[2..nimp], old[2..nold] of its time array
"""
import numpy as np

FIELD_ORDER = ["w", "z", "p", "s", "xi", "phi", "b", "aj", "b_ic", "aj_ic"]
_PARAMS = ["ra", "pr", "raxi", "sc", "prmag", "ek", "stef", "radratio", "sigma_ratio"]
_TRUNC = ["n_r_max", "n_theta_max", "n_phi_tot", "minc", "nalias", "n_r_ic_max"]
_ROT = ["omega_ic1", "omegaOsz_ic1", "tOmega_ic1", "omega_ic2", "omegaOsz_ic2", "tOmega_ic2",
        "omega_ma1", "omegaOsz_ma1", "tOmega_ma1", "omega_ma2", "omegaOsz_ma2", "tOmega_ma2"]


class CheckpointError(RuntimeError):
    pass


class Checkpoint:
    """The content of one checkpoint file.  `fields[name]` is complex128 [n_r, lm_max] (n_r = n_r_ic_max for *_ic);
    `past[name]` = dict(expl=[...], impl=[...], old=[...]) holds the older levels 2.. of the field's time array
    (`type_tarray`, time_array.f90); `scalars_past` the same for domega_ic_dt / domega_ma_dt (and the Lorentz torques of
    files older than version 5)."""

    def __init__(self):
        self.version = 5
        self.time = 0.0
        self.family = "MULTISTEP"
        self.nexp, self.nimp, self.nold = 2, 1, 1
        self.dt = np.zeros(2)
        self.n_time_step = 0
        self.params = {k: 0.0 for k in _PARAMS}
        self.trunc = {k: 0 for k in _TRUNC}
        self.l_max = self.m_min = self.m_max = 0
        self.rscheme = dict(version="cheb", n=0, n2=0, ratio1=0.0, ratio2=0.0)
        self.r = np.zeros(0)
        self.scalars_past = {}
        self.rotation = {k: 0.0 for k in _ROT}
        self.fields = {}
        self.past = {}

    @property
    def lm_max(self):
        minc = self.trunc["minc"]
        return sum(self.l_max - m + 1 for m in range(self.m_min, self.m_max + 1, minc))

    def lm_maps(self):
        """st_map (blocking.f90:293-337): lm2l, lm2m with m outer, l inner."""
        minc = self.trunc["minc"]
        l, m = [], []
        for mm in range(self.m_min, self.m_max + 1, minc):
            for ll in range(mm, self.l_max + 1):
                l.append(ll)
                m.append(mm)
        return np.array(l, dtype=np.int32), np.array(m, dtype=np.int32)

    def _levels(self):
        if not self.family.startswith("MULTISTEP"):
            return (("expl", 0), ("impl", 0), ("old", 0))
        return (("expl", self.nexp - 1), ("impl", self.nimp - 1), ("old", self.nold - 1))


def _take(f, dtype, n):
    a = np.fromfile(f, dtype=dtype, count=n)
    if a.size != n:
        raise CheckpointError("checkpoint file ends early")
    return a


def read_checkpoint(path):
    """readStartFields (readCheckPoints.f90:840-1060, :1508-1601) without any remapping."""
    ck = Checkpoint()
    with open(path, "rb") as f:
        ck.version = int(_take(f, "<i4", 1)[0])
        if not 2 <= ck.version <= 5:
            # version 1 is the pre-time-array layout (readCheckPoints.f90:878-889); larger values mean record markers or
            # the other endianness (readCheckPoints.f90:844-870)
            raise CheckpointError(f"unsupported checkpoint version {ck.version}")
        ck.time = float(_take(f, "<f8", 1)[0])
        ck.family = f.read(10).decode("ascii").rstrip()
        ck.nexp, ck.nimp, ck.nold = (int(x) for x in _take(f, "<i4", 3))
        multistep = ck.family.startswith("MULTISTEP")
        if not multistep and not ck.family.startswith("DIRK"):
            raise CheckpointError(f"unknown time-scheme family '{ck.family}'")
        ck.dt = _take(f, "<f8", ck.nexp if multistep else 1).copy()
        ck.n_time_step = int(_take(f, "<i4", 1)[0])
        names = [k for k in _PARAMS if not (k == "stef" and ck.version <= 2)]
        ck.params.update(zip(names, (float(x) for x in _take(f, "<f8", len(names)))))
        ck.trunc = dict(zip(_TRUNC, (int(x) for x in _take(f, "<i4", 6))))
        if ck.version > 3:
            ck.l_max, ck.m_min, ck.m_max = (int(x) for x in _take(f, "<i4", 3))
        else:   # readCheckPoints.f90:926-937
            t = ck.trunc
            ck.l_max = t["nalias"] * t["n_theta_max"] // 30 if t["n_phi_tot"] == 1 else t["nalias"] * t["n_phi_tot"] // 60
            ck.m_min = 0
            ck.m_max = 0 if t["n_phi_tot"] == 1 else (ck.l_max // t["minc"]) * t["minc"]
        ck.rscheme["version"] = f.read(72).decode("ascii").rstrip()
        ck.rscheme["n"], ck.rscheme["n2"] = (int(x) for x in _take(f, "<i4", 2))
        ck.rscheme["ratio1"], ck.rscheme["ratio2"] = (float(x) for x in _take(f, "<f8", 2))
        n_r, n_ic, lm_max = ck.trunc["n_r_max"], ck.trunc["n_r_ic_max"], ck.lm_max
        ck.r = _take(f, "<f8", n_r).copy()
        if multistep:
            scal = ["domega_ic_dt", "domega_ma_dt"] + (["lorentz_torque_ic_dt", "lorentz_torque_ma_dt"] if ck.version < 5 else [])
            for nm in scal:
                ck.scalars_past[nm] = {k: _take(f, "<f8", n).copy() for k, n in ck._levels()}
        ck.rotation = dict(zip(_ROT, (float(x) for x in _take(f, "<f8", 12))))
        flags = ["l_heat", "l_chemical_conv", "l_phase_field", "l_mag", "l_press_store", "l_cond_ic"]
        if ck.version <= 2:
            flags.remove("l_phase_field")
        fl = dict(zip(flags, (bool(x) for x in _take(f, "<i4", len(flags)))))
        present = {"w": True, "z": True, "p": fl["l_press_store"], "s": fl["l_heat"], "xi": fl["l_chemical_conv"],
                   "phi": fl.get("l_phase_field", False), "b": fl["l_mag"], "aj": fl["l_mag"],
                   "b_ic": fl["l_mag"] and fl["l_cond_ic"], "aj_ic": fl["l_mag"] and fl["l_cond_ic"]}
        for nm in FIELD_ORDER:
            if not present[nm]:
                continue
            rows = n_ic if nm.endswith("_ic") else n_r
            ck.fields[nm] = _take(f, "<c16", rows * lm_max).reshape(rows, lm_max).copy()
            ck.past[nm] = {k: [_take(f, "<c16", rows * lm_max).reshape(rows, lm_max).copy() for _ in range(n)]
                           for k, n in ck._levels()}
        if f.read(1):
            raise CheckpointError("trailing bytes after the last field: truncation or flags do not match the file")
    return ck


def write_checkpoint(path, ck):
    """store (storeCheckPoints.f90:45-277): always the current layout, version 5."""
    multistep = ck.family.startswith("MULTISTEP")
    n_r, n_ic, lm_max = ck.trunc["n_r_max"], ck.trunc["n_r_ic_max"], ck.lm_max
    if len(ck.r) != n_r:
        raise CheckpointError("radius array does not have n_r_max entries")

    def levels_of(store, nm, dtype, shape):
        out = []
        for k, n in ck._levels():
            have = store.get(nm, {}).get(k, [])
            if len(have) != n:
                raise CheckpointError(f"time array '{nm}' needs {n} past '{k}' level(s), got {len(have)}")
            for a in have:
                a = np.ascontiguousarray(a, dtype=dtype)
                if a.shape != shape:
                    raise CheckpointError(f"time array '{nm}' has shape {a.shape}, expected {shape}")
                out.append(a)
        return out

    with open(path, "wb") as f:
        np.array([5], "<i4").tofile(f)
        np.array([ck.time], "<f8").tofile(f)
        f.write(ck.family.ljust(10).encode("ascii")[:10])
        np.array([ck.nexp, ck.nimp, ck.nold], "<i4").tofile(f)
        dt = np.atleast_1d(np.asarray(ck.dt, dtype="<f8"))
        if dt.size != (ck.nexp if multistep else 1):
            raise CheckpointError("dt array does not match the time scheme")
        dt.tofile(f)
        np.array([ck.n_time_step], "<i4").tofile(f)
        np.array([ck.params[k] for k in _PARAMS], "<f8").tofile(f)
        np.array([ck.trunc[k] for k in _TRUNC], "<i4").tofile(f)
        np.array([ck.l_max, ck.m_min, ck.m_max], "<i4").tofile(f)
        f.write(ck.rscheme["version"].ljust(72).encode("ascii")[:72])
        np.array([ck.rscheme["n"], ck.rscheme["n2"]], "<i4").tofile(f)
        np.array([ck.rscheme["ratio1"], ck.rscheme["ratio2"]], "<f8").tofile(f)
        np.asarray(ck.r, dtype="<f8").tofile(f)
        if multistep:
            for nm in ("domega_ic_dt", "domega_ma_dt"):
                for k, n in ck._levels():
                    a = np.asarray(ck.scalars_past.get(nm, {}).get(k, np.zeros(n)), dtype="<f8")
                    if a.size != n:
                        raise CheckpointError(f"scalar time array '{nm}' needs {n} past '{k}' level(s)")
                    a.tofile(f)
        np.array([ck.rotation[k] for k in _ROT], "<f8").tofile(f)
        has = ck.fields
        if "w" not in has or "z" not in has or (("b" in has) != ("aj" in has)) or (("b_ic" in has) != ("aj_ic" in has)):
            raise CheckpointError("a checkpoint holds at least w and z; b/aj and b_ic/aj_ic come in pairs")
        np.array([int("s" in has), int("xi" in has), int("phi" in has), int("b" in has), int("p" in has), int("b_ic" in has)],
                 "<i4").tofile(f)
        for nm in FIELD_ORDER:
            if nm not in has:
                continue
            shape = (n_ic if nm.endswith("_ic") else n_r, lm_max)
            a = np.ascontiguousarray(has[nm], dtype="<c16")
            if a.shape != shape:
                raise CheckpointError(f"field '{nm}' has shape {a.shape}, expected {shape}")
            a.tofile(f)
            for lev in levels_of(ck.past, nm, "<c16", shape):
                lev.tofile(f)
