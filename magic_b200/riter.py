"""Host-side mirror of `rIter_single_t%radialLoop` (rIter.f90:94-464): the batched radial loop.

`RadialLoop.radialLoop(fields)` takes the R-distributed spectral containers of one rank and returns the
explicit terms (dwdt, dzdt, dpdt, dsdt, dbdt, djdt, dVxVhLM, dVxBhLM, dVSrLM, dtrkc, dthkc) exactly like the
Fortran subroutine, for all local radial levels at once.
"""
import ctypes as C
from ctypes import byref, c_double, c_int, c_void_p

import numpy as np

from .lib import MagicError, check, load_library, ptr


class Params(C.Structure):
    """magic_params (include/magic_sht.h): logic.f90 / physical_parameters.f90 values the loop reads."""
    _ints = ["l_conv", "l_mag", "l_heat", "l_conv_nl", "l_heat_nl", "l_mag_nl", "l_mag_LF", "l_mag_kin", "l_anel",
             "l_adv_curl", "l_corr", "l_double_curl", "l_single_matrix", "l_chemical_conv", "l_precession",
             "l_centrifuge", "l_anelastic_liquid", "l_cour_alf_damp", "l_full_sphere", "l_parallel_solve",
             "l_temperature_diff", "ktopv", "kbotv", "l_cond_ma", "l_cond_ic", "l_rot_ma", "l_rot_ic", "n_r_max",
             "n_r_LCR"]
    _dbls = ["LFfac", "CorFac", "epsc", "epscXi", "opm", "ViscHeatFac", "OhmLossFac", "oek", "po", "prec_angle",
             "dilution_fac", "ra", "opr", "omega_ma", "omega_ic", "r_cmb", "r_icb", "courfac", "alffac",
             "epsPhase", "phaseDiffFac", "penaltyFac", "tmelt"]
    _fields_ = [(n, c_int) for n in _ints] + [(n, c_double) for n in _dbls] + [("l_phase_field", c_int)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


RADIAL_NAMES = ["r", "or1", "or2", "or4", "orho1", "orho2", "beta", "rho0", "otemp1", "temp0", "visc", "lambda",
                "epscProf", "delxr2", "delxh2"]
# magic_rloop_diagnostics: mask bits and slots (include/magic_sht.h)
DIAG_HEL, DIAG_HEMI, DIAG_POWER, DIAG_PERPPAR, DIAG_FLUX, DIAG_VISCBC, DIAG_PHASE, DIAG_RMSBULK = 1, 2, 4, 8, 16, 32, 64, 256
NDIAG = 40
NTO = 15
NRMS = 14
IN_NAMES = ["w", "dw", "ddw", "z", "dz", "s", "ds", "p", "xi", "b", "db", "ddb", "aj", "dj", "phi"]
OUT_NAMES = ["dwdt", "dzdt", "dpdt", "dsdt", "dxidt", "dbdt", "djdt", "dVxVhLM", "dVxBhLM", "dVSrLM", "dVXirLM", "dphidt"]


class _Radial(C.Structure):
    _fields_ = [("nR", c_void_p), ("l_R", c_void_p)] + [(n + "_", c_void_p) for n in RADIAL_NAMES]


class _FieldsIn(C.Structure):
    _fields_ = [(n, c_void_p) for n in IN_NAMES]


class _FieldsOut(C.Structure):
    _fields_ = [(n, c_void_p) for n in OUT_NAMES[:11]] + [("dtrkc", c_void_p), ("dthkc", c_void_p), ("dphidt", c_void_p)]


def level_chunks(n_r_loc, level_chunk):
    """magic_level_chunks: (start, size) lists of the level chunks of a slab (host only)."""
    lib = load_library()
    n = c_int()
    st = (c_int * n_r_loc)()
    sz = (c_int * n_r_loc)()
    check(lib.magic_level_chunks(c_int(n_r_loc), c_int(level_chunk), byref(n), st, sz))
    return list(st[:n.value]), list(sz[:n.value])


class _LmIn(C.Structure):
    _fields_ = [("flow", c_void_p), ("s", c_void_p), ("field", c_void_p), ("xi", c_void_p)]


class _LmOut(C.Structure):
    _fields_ = [("dflowdt", c_void_p), ("dsdt", c_void_p), ("dbdt", c_void_p), ("dtrkc", c_void_p), ("dthkc", c_void_p),
                ("dxidt", c_void_p)]


class RadialLoop:
    """initialize_radialLoop + radialLoopG (radialLoop.f90:26-101) for the levels of one rank."""

    def __init__(self, sht, params, radial, level_chunk=0):
        """radial: dict with 'nR' (global 1-based level numbers), 'l_R' and the arrays of RADIAL_NAMES,
        one entry per local level (radial_functions, radial.f90:283-307)."""
        self.lib = load_library()
        self.sht = sht
        self.params = params
        self.n_r_loc = len(radial["nR"])
        keep = []
        rad = _Radial()
        for nm in ["nR", "l_R"]:
            a = np.ascontiguousarray(radial[nm], dtype=np.int32)
            keep.append(a)
            setattr(rad, nm, a.ctypes.data)
        for nm in RADIAL_NAMES:
            a = np.ascontiguousarray(radial.get(nm, np.ones(self.n_r_loc)), dtype=np.float64)
            if a.shape != (self.n_r_loc,):
                raise ValueError(f"radial['{nm}'] must have {self.n_r_loc} entries")
            keep.append(a)
            setattr(rad, nm + "_", a.ctypes.data)
        self._h = c_void_p()
        check(self.lib.magic_rloop_create(sht.handle, byref(params), byref(rad), c_int(self.n_r_loc), c_int(level_chunk),
                                          byref(self._h)))

    def finalize(self):
        if self._h:
            self.lib.magic_rloop_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.finalize()
        except Exception:
            pass

    def _structs(self, fields, outs, dtrkc, dthkc, device):
        fin = _FieldsIn()
        fout = _FieldsOut()
        keep = []
        for nm in IN_NAMES:
            v = fields.get(nm)
            if v is None:
                continue
            if device:
                setattr(fin, nm, int(v))
            else:
                a = np.ascontiguousarray(v, dtype=np.complex128)
                if a.shape != (self.n_r_loc, self.sht.lm_max):
                    raise ValueError(f"field '{nm}' must have shape ({self.n_r_loc}, {self.sht.lm_max})")
                keep.append(a)
                setattr(fin, nm, a.ctypes.data)
        for nm in OUT_NAMES:
            v = outs.get(nm)
            if v is None:
                continue
            setattr(fout, nm, int(v) if device else v.ctypes.data)
        fout.dtrkc = int(dtrkc) if device else dtrkc.ctypes.data
        fout.dthkc = int(dthkc) if device else dthkc.ctypes.data
        return fin, fout, keep

    def radialLoop(self, fields, time=0.0, out=None):
        """Host-buffer call (what the Fortran rIter_cuda_t does).  fields: name -> complex128 [n_r_loc, lm_max].
        Returns dict of outputs; pass `out` to reuse arrays."""
        if out is None:
            out = {nm: np.zeros((self.n_r_loc, self.sht.lm_max), dtype=np.complex128) for nm in OUT_NAMES}
            out["dtrkc"] = np.zeros(self.n_r_loc)
            out["dthkc"] = np.zeros(self.n_r_loc)
        fin, fout, keep = self._structs(fields, out, out["dtrkc"], out["dthkc"], device=False)
        check(self.lib.magic_rloop_run(self._h, byref(fin), byref(fout), c_double(time)))
        return out

    def diagnostics(self, fields, mask, ktops=1, kbots=1, device=False):
        """The in-loop diagnostics of a log step (rIter.f90:303-373: get_helicity, get_hemi, get_visc_heat, get_perpPar,
        get_fluxes, get_nlBLayers) for this rank's levels: float64 [n_r_loc, NDIAG], slots as documented in
        include/magic_sht.h.  fields as for radialLoop (host arrays), or device pointers with device=True."""
        fin, _, keep = self._structs(fields, {}, 0 if device else np.zeros(1), 0 if device else np.zeros(1), device=device)
        out = np.zeros((self.n_r_loc, NDIAG))
        fn = self.lib.magic_rloop_diagnostics_dev if device else self.lib.magic_rloop_diagnostics
        check(fn(self._h, byref(fin), c_int(mask), c_int(ktops), c_int(kbots), out.ctypes.data_as(c_void_p)))
        return out

    @staticmethod
    def _result(out, shape):
        if out is None:
            return np.zeros(shape, dtype=np.complex128)
        if out.shape != shape or out.dtype != np.complex128 or not out.flags.c_contiguous:
            raise ValueError(f"out must be a C-contiguous complex128 array of shape {shape}")
        return out

    def dtb(self, fields, device=False, out=None):
        """get_dtBLM (dtB.f90:144-223) for this rank's levels: complex128 [11, n_r_loc, lm_max] (BtVrLM, BpVrLM, BrVtLM, BrVpLM,
        BtVpLM, BpVtLM, BpVtBtVpCotLM, BpVtBtVpSn2LM, BrVZLM, BtVZLM, BtVZsn2LM).  `out`: a host array to fill."""
        fin, _, keep = self._structs(fields, {}, 0 if device else np.zeros(1), 0 if device else np.zeros(1), device=device)
        out = self._result(out, (11, self.n_r_loc, self.sht.lm_max))
        fn = self.lib.magic_rloop_dtb_dev if device else self.lib.magic_rloop_dtb
        check(fn(self._h, byref(fin), out.ctypes.data_as(c_void_p)))
        return out

    def to_next(self, fields, device=False):
        """getTOnext's grid part (rIter.f90:395-398, TO.f90:330-343): keeps Bs, Bp, Bz of every local level on the device."""
        fin, _, keep = self._structs(fields, {}, 0 if device else np.zeros(1), 0 if device else np.zeros(1), device=device)
        fn = self.lib.magic_rloop_to_next_dev if device else self.lib.magic_rloop_to_next
        check(fn(self._h, byref(fin)))

    def to(self, fields, dtLast, device=False):
        """getTO (rIter.f90:400-404, TO.f90:141-307) for this rank's levels: float64 [n_r_loc, NTO, n_theta_max], colatitudes north ->
        south, arrays as documented in include/magic_sht.h."""
        fin, _, keep = self._structs(fields, {}, 0 if device else np.zeros(1), 0 if device else np.zeros(1), device=device)
        out = np.zeros((self.n_r_loc, NTO, self.sht.n_theta_max))
        fn = self.lib.magic_rloop_to_dev if device else self.lib.magic_rloop_to
        check(fn(self._h, byref(fin), c_double(dtLast), out.ctypes.data_as(c_void_p)))
        return out

    def rms_keep(self, fields, device=False):
        """Keeps w, dw, z of every local level on the device: the previous step's velocity of get_nl_RMS (RMS.f90:545-551)."""
        fin, _, keep = self._structs(fields, {}, 0 if device else np.zeros(1), 0 if device else np.zeros(1), device=device)
        fn = self.lib.magic_rloop_rms_keep_dev if device else self.lib.magic_rloop_rms_keep
        check(fn(self._h, byref(fin)))

    def rms(self, fields, dt, device=False, out=None):
        """The r.m.s. force balance inside the radial loop on lRmsCalc steps (rIter.f90:215-252, 710; RMS.f90:469-610) for this
        rank's levels: complex128 [NRMS, n_r_loc, lm_max], arrays as documented in include/magic_sht.h.  `out`: a host array
        to fill (page-locked memory takes the result at PCIe speed)."""
        fin, _, keep = self._structs(fields, {}, 0 if device else np.zeros(1), 0 if device else np.zeros(1), device=device)
        out = self._result(out, (NRMS, self.n_r_loc, self.sht.lm_max))
        fn = self.lib.magic_rloop_rms_dev if device else self.lib.magic_rloop_rms
        check(fn(self._h, byref(fin), c_double(dt), out.ctypes.data_as(c_void_p)))
        return out

    def graph_fields(self, fields, level, mag=False, pressure=False):
        """The grid fields graphOut_mpi reads (rIter.f90:303-314) for local level `level`: dict of float64
        [n_phi, nlat_padded] arrays (Fortran f(nlat_padded, n_phi))."""
        fin, _, keep = self._structs(fields, {}, np.zeros(1), np.zeros(1), device=False)
        shp = (self.sht.n_phi_max, self.sht.nlat_padded)
        names = ["vr", "vt", "vp"] + (["br", "bt", "bp"] if mag else []) + ["sr"] + (["pr"] if pressure else [])
        out = {n: np.zeros(shp) for n in names}
        ptrs = [out[n].ctypes.data_as(c_void_p) if n in out else c_void_p(None) for n in ("vr", "vt", "vp", "br", "bt", "bp", "sr", "pr")]
        check(self.lib.magic_rloop_graph_fields(self._h, byref(fin), c_int(level), *ptrs))
        return out

    def radialLoop_dev(self, fields_dev, out_dev, dtrkc_dev, dthkc_dev, time=0.0):
        """Device-pointer call: dict name -> int device pointer (e.g. torch tensor .data_ptr())."""
        fin, fout, keep = self._structs(fields_dev, out_dev, dtrkc_dev, dthkc_dev, device=True)
        check(self.lib.magic_rloop_run_dev(self._h, byref(fin), byref(fout), c_double(time)))

    def run_lm_dev(self, transposer, flow_LM, s_LM, field_LM, dflowdt_LM, dsdt_LM, dbdt_LM, dtrkc_dev, dthkc_dev, time=0.0):
        """The whole hot path of a step on LM-distributed device containers (ints = device pointers; field/dbdt 0 without
        l_mag): transp_LMloc_to_Rloc -> radialLoopG -> transp_Rloc_to_LMloc (step_time.f90:485-612), the transposes
        pipelined chunk by chunk against the compute when there is more than one rank."""
        i = _LmIn(int(flow_LM), int(s_LM) or None, int(field_LM) or None, None)
        o = _LmOut(int(dflowdt_LM), int(dsdt_LM) or None, int(dbdt_LM) or None, int(dtrkc_dev), int(dthkc_dev), None)
        check(self.lib.magic_rloop_run_lm_dev(self._h, transposer._h, byref(i), byref(o), c_double(time)))

    def run_lm(self, transposer, lm_in, lm_out, dtrkc, dthkc, time=0.0):
        """magic_rloop_run_lm: the same on HOST LM-distributed containers (numpy complex128 arrays [nf, n_r_max, nlm_loc];
        dict keys flow, s, field, xi / dflowdt, dsdt, dbdt, dxidt; dtrkc, dthkc float64 [n_r_loc])."""
        def a(d, k):
            v = d.get(k)
            return None if v is None else v.ctypes.data
        i = _LmIn(a(lm_in, "flow"), a(lm_in, "s"), a(lm_in, "field"), a(lm_in, "xi"))
        o = _LmOut(a(lm_out, "dflowdt"), a(lm_out, "dsdt"), a(lm_out, "dbdt"), dtrkc.ctypes.data, dthkc.ctypes.data, a(lm_out, "dxidt"))
        check(self.lib.magic_rloop_run_lm(self._h, transposer._h, byref(i), byref(o), c_double(time)))

    def set_radial_matrices(self, D1, D2):
        """Dense matrices of the host's radial scheme (get_dr / get_ddr incl. the n_cheb_max truncation), [n_r_max, n_r_max]."""
        D1 = np.ascontiguousarray(D1, dtype=np.float64)
        D2 = np.ascontiguousarray(D2, dtype=np.float64)
        check(self.lib.magic_rloop_set_radial_matrices(self._h, c_int(D1.shape[0]), c_void_p(D1.ctypes.data), c_void_p(D2.ctypes.data)))

    def set_lm_radial(self, or2, orho1, dentropy0, l_R):
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (or2, orho1, dentropy0)]
        lr = np.ascontiguousarray(l_R, dtype=np.int32)
        check(self.lib.magic_rloop_set_lm_radial(self._h, c_int(len(lr)), *[c_void_p(x.ctypes.data) for x in a], c_void_p(lr.ctypes.data)))

    def lm_options(self, derivs_on_device=False, finish_on_device=False):
        check(self.lib.magic_rloop_lm_options(self._h, c_int(int(derivs_on_device)), c_int(int(finish_on_device))))

    def set_rotation(self, omega_ma, omega_ic):
        """Boundary rotation rates of the coming step (omega_ma, omega_ic of v_rigid_boundary)."""
        check(self.lib.magic_rloop_set_rotation(self._h, c_double(omega_ma), c_double(omega_ic)))

    def torques(self):
        """(lorentz_torque_ic, lorentz_torque_ma) of the last run (rIter.f90:279-292,461)."""
        a, b = c_double(), c_double()
        check(self.lib.magic_rloop_get_torques(self._h, byref(a), byref(b)))
        return a.value, b.value

    def br_v_bcs(self, boundary):
        """(br_vt_lm, br_vp_lm) of get_br_v_bcs (nonlinear_bcs.f90:24-74) for boundary 'CMB' or 'ICB' after a run."""
        n = self.sht.lm_max
        a, b = np.zeros(n, dtype=np.complex128), np.zeros(n, dtype=np.complex128)
        check(self.lib.magic_rloop_get_br_v_bcs(self._h, c_int({"CMB": 0, "ICB": 1}[boundary]), a.ctypes.data_as(c_void_p),
                                                b.ctypes.data_as(c_void_p)))
        return a, b

    def sync(self):
        check(self.lib.magic_rloop_sync(self._h))

    def launch_count(self):
        return int(self.lib.magic_rloop_launch_count(self._h))

    def last_timing(self):
        t = (c_double * 8)()
        check(self.lib.magic_rloop_last_timing(self._h, t))
        keys = ["total", "prep", "legendre_syn", "fft_c2r", "get_nl", "fft_r2c", "legendre_an", "get_td"]
        return dict(zip(keys, list(t)))

    def last_exposed(self):
        """(ms before the first chunk's first kernel, ms after the last chunk's last kernel) of the last run."""
        t = (c_double * 2)()
        check(self.lib.magic_rloop_last_exposed(self._h, t))
        return float(t[0]), float(t[1])

    def legendre_flops(self):
        return float(self.lib.magic_rloop_legendre_flops(self._h))

    def level_chunk(self):
        return int(self.lib.magic_rloop_level_chunk(self._h))

    def pin_host(self, array):
        """magic_rloop_pin_host: page-lock a persistent numpy array of the caller (unpin_host before it is freed)."""
        check(self.lib.magic_rloop_pin_host(self._h, c_void_p(array.ctypes.data), array.nbytes))

    def unpin_host(self, array):
        check(self.lib.magic_rloop_unpin_host(self._h, c_void_p(array.ctypes.data)))

    def legendre_units(self):
        """(reference count, executed) scalar-equivalent Legendre passes per bulk level."""
        u = (c_double * 2)()
        check(self.lib.magic_rloop_legendre_units(self._h, u))
        return float(u[0]), float(u[1])
