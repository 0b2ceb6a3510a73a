"""ctypes front-end to the CPU oracle (oracle/magic_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; magic_b200/ never does.  Parity status (what is pinned to reference vectors, what is not): see
oracle/magic_oracle.h.

Array conventions (numpy, C order):
  spectral  complex128 [lm_max]            st_map order (reference X(lm))
  grid      float64    [n_phi, n_theta]    theta fastest  (reference f(nlat_padded, n_phi_max)),
                                           theta rows N/S interleaved like the native backend
"""
import ctypes as C
import os
import subprocess
from ctypes import POINTER, c_double, c_int, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(force=False):
    """Compile the oracle with gcc (oracle/Makefile)."""
    if force or not (os.path.exists(os.path.join(_HERE, "libmagic_oracle.so"))
                     and os.path.exists(os.path.join(_HERE, "libmagic_oracle_fast.so"))):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []),
                              stdout=subprocess.DEVNULL)


class Params(C.Structure):
    """orc_params: logic.f90 / physical_parameters.f90 values read by the hot path."""
    _ints = ["l_conv", "l_mag", "l_heat", "l_conv_nl", "l_heat_nl", "l_mag_nl", "l_mag_LF", "l_mag_kin",
             "l_anel", "l_adv_curl", "l_corr", "l_double_curl", "l_single_matrix", "l_chemical_conv",
             "l_precession", "l_centrifuge", "l_anelastic_liquid", "l_cour_alf_damp", "l_full_sphere",
             "l_parallel_solve", "l_temperature_diff", "ktopv", "kbotv", "l_cond_ma", "l_cond_ic",
             "l_rot_ma", "l_rot_ic", "n_r_max", "n_r_LCR"]
    _dbls = ["LFfac", "CorFac", "epsc", "epscXi", "opm", "ViscHeatFac", "OhmLossFac", "oek", "po",
             "prec_angle", "dilution_fac", "ra", "opr", "omega_ma", "omega_ic", "r_cmb", "r_icb",
             "courfac", "alffac", "epsPhase", "phaseDiffFac", "penaltyFac", "tmelt"]
    _fields_ = [(n, c_int) for n in _ints] + [(n, c_double) for n in _dbls] + [("l_phase_field", c_int)]


_RAD_NAMES = ["r", "or1", "or2", "or4", "orho1", "orho2", "beta", "rho0", "otemp1", "temp0", "visc",
              "lambda_", "epscProf", "delxr2", "delxh2"]
_IN_NAMES = ["w", "dw", "ddw", "z", "dz", "s", "ds", "p", "xi", "b", "db", "ddb", "aj", "dj", "phi"]
_OUT_NAMES = ["dwdt", "dzdt", "dpdt", "dsdt", "dxidt", "dbdt", "djdt", "dVxVhLM", "dVxBhLM", "dVSrLM",
              "dVXirLM"]


class _Radial(C.Structure):
    _fields_ = [("nR", c_void_p), ("l_R", c_void_p)] + [(n, c_void_p) for n in _RAD_NAMES]


class _FieldsIn(C.Structure):
    _fields_ = [(n, c_void_p) for n in _IN_NAMES]


class _FieldsOut(C.Structure):
    _fields_ = [(n, c_void_p) for n in _OUT_NAMES] + [("dtrkc", c_void_p), ("dthkc", c_void_p),
                                                       ("lorentz_torque_ic", c_void_p), ("lorentz_torque_ma", c_void_p),
                                                       ("br_vt_lm_cmb", c_void_p), ("br_vp_lm_cmb", c_void_p),
                                                       ("br_vt_lm_icb", c_void_p), ("br_vp_lm_icb", c_void_p), ("dphidt", c_void_p)]


def _load(fast):
    build()
    lib = C.CDLL(os.path.join(_HERE, "libmagic_oracle_fast.so" if fast else "libmagic_oracle.so"))
    lib.orc_create.restype = c_void_p
    lib.orc_create.argtypes = [c_int] * 5
    for name in ["orc_lm2l", "orc_lm2m", "orc_lm2lmS", "orc_lm2lmA"]:
        getattr(lib, name).restype = POINTER(c_int)
        getattr(lib, name).argtypes = [c_void_p]
    for name in ["orc_theta_ord", "orc_gauss", "orc_plm", "orc_dplm"]:
        getattr(lib, name).restype = POINTER(c_double)
        getattr(lib, name).argtypes = [c_void_p]
    lib.orc_theta_vec.restype = POINTER(c_double)
    lib.orc_theta_vec.argtypes = [c_void_p, c_int]
    lib.orc_lm_vec.restype = POINTER(c_double)
    lib.orc_lm_vec.argtypes = [c_void_p, c_int]
    return lib


def grid_sizes(l_max=0, n_phi_tot=0, minc=1, nalias=20):
    """truncation.f90:55-105 -> dict(l_max, m_max, n_theta_max, n_phi_max, n_m_max, lm_max, n_phi_tot)."""
    lib = _load(False)
    out = (c_int * 7)()
    lib.orc_grid_sizes(c_int(l_max), c_int(n_phi_tot), c_int(minc), c_int(nalias), out)
    keys = ["l_max", "m_max", "n_theta_max", "n_phi_max", "n_m_max", "lm_max", "n_phi_tot"]
    return dict(zip(keys, list(out)))


def get_blocks(n_points, n_procs):
    """parallel.f90:75-92 getBlocks -> (start[], stop[]) 1-based inclusive."""
    lib = _load(False)
    s = (c_int * n_procs)()
    e = (c_int * n_procs)()
    lib.orc_get_blocks(c_int(n_points), c_int(n_procs), s, e)
    return np.array(s), np.array(e)


def _p(a):
    return a.ctypes.data_as(c_void_p) if a is not None else None


class Oracle:
    """One `module sht` + radial-loop context of the reference's native backend."""

    def __init__(self, l_max, minc=1, n_theta=None, n_phi=None, m_max=None, nalias=20, fast=False, threads=1):
        self.lib = _load(fast)
        if n_theta is None:
            gs = grid_sizes(l_max=l_max, minc=minc, nalias=nalias)
            n_theta, n_phi = gs["n_theta_max"], gs["n_phi_max"]
        if m_max is None:
            m_max = (l_max // minc) * minc
        self.l_max, self.m_max, self.minc, self.n_theta, self.n_phi = l_max, m_max, minc, n_theta, n_phi
        self.h = c_void_p(self.lib.orc_create(l_max, m_max, minc, n_theta, n_phi))
        self.lib.orc_set_threads(self.h, c_int(threads))
        self.lm_max = self.lib.orc_lm_max(self.h)
        self.n_m_max = self.lib.orc_n_m_max(self.h)
        n = self.lm_max
        self.lm2l = np.ctypeslib.as_array(self.lib.orc_lm2l(self.h), (n,)).copy()
        self.lm2m = np.ctypeslib.as_array(self.lib.orc_lm2m(self.h), (n,)).copy()
        self.lm2lmS = np.ctypeslib.as_array(self.lib.orc_lm2lmS(self.h), (n,)).copy()
        self.lm2lmA = np.ctypeslib.as_array(self.lib.orc_lm2lmA(self.h), (n,)).copy()
        self.theta_ord = np.ctypeslib.as_array(self.lib.orc_theta_ord(self.h), (n_theta,)).copy()
        self.gauss = np.ctypeslib.as_array(self.lib.orc_gauss(self.h), (n_theta,)).copy()
        names = ["sinTheta", "cosTheta", "O_sin_theta", "O_sin_theta_E2", "sinTheta_E2", "cosn_theta_E2"]
        for i, nm in enumerate(names):
            setattr(self, nm, np.ctypeslib.as_array(self.lib.orc_theta_vec(self.h, i), (n_theta,)).copy())
        names = ["dLh", "dTheta1S", "dTheta1A", "dTheta2S", "dTheta2A", "dTheta3S", "dTheta3A", "dTheta4S",
                 "dTheta4A"]
        for i, nm in enumerate(names):
            setattr(self, nm, np.ctypeslib.as_array(self.lib.orc_lm_vec(self.h, i), (n,)).copy())

    def __del__(self):
        try:
            self.lib.orc_destroy(self.h)
        except Exception:
            pass

    def set_threads(self, n):
        self.lib.orc_set_threads(self.h, c_int(n))

    # tables ------------------------------------------------------------------------------------
    def plm(self):
        return np.ctypeslib.as_array(self.lib.orc_plm(self.h), (self.n_theta // 2, self.lm_max)).copy()

    def dplm(self):
        return np.ctypeslib.as_array(self.lib.orc_dplm(self.h), (self.n_theta // 2, self.lm_max)).copy()

    def lo_map(self, n_procs):
        lo2st = np.zeros(self.lm_max, dtype=np.int32)
        s = np.zeros(n_procs, dtype=np.int32)
        e = np.zeros(n_procs, dtype=np.int32)
        self.lib.orc_lo_map(self.h, c_int(n_procs), _p(lo2st), _p(s), _p(e))
        return lo2st, s, e

    # helpers -----------------------------------------------------------------------------------
    def _g(self):
        return np.zeros((self.n_phi, self.n_theta), dtype=np.float64)

    def _s(self):
        return np.zeros(self.lm_max, dtype=np.complex128)

    @staticmethod
    def _c(a):
        return np.ascontiguousarray(a, dtype=np.complex128)

    @staticmethod
    def _r(a):
        return np.ascontiguousarray(a, dtype=np.float64)

    # module sht (sht_native.f90:16-20) ---------------------------------------------------------
    def scal_to_spat(self, Slm, lcut):
        f = self._g()
        self.lib.orc_scal_to_spat(self.h, _p(self._c(Slm)), _p(f), c_int(lcut))
        return f

    def scal_to_grad_spat(self, Slm, lcut):
        a, b = self._g(), self._g()
        self.lib.orc_scal_to_grad_spat(self.h, _p(self._c(Slm)), _p(a), _p(b), c_int(lcut))
        return a, b

    def pol_to_grad_spat(self, Slm, lcut):
        a, b = self._g(), self._g()
        self.lib.orc_pol_to_grad_spat(self.h, _p(self._c(Slm)), _p(a), _p(b), c_int(lcut))
        return a, b

    def torpol_to_spat(self, W, dW, Z, lcut):
        a, b, c = self._g(), self._g(), self._g()
        self.lib.orc_torpol_to_spat(self.h, _p(self._c(W)), _p(self._c(dW)), _p(self._c(Z)), _p(a), _p(b), _p(c),
                                    c_int(lcut))
        return a, b, c

    def sphtor_to_spat(self, dW, Z, lcut):
        a, b = self._g(), self._g()
        self.lib.orc_sphtor_to_spat(self.h, _p(self._c(dW)), _p(self._c(Z)), _p(a), _p(b), c_int(lcut))
        return a, b

    def torpol_to_dphspat(self, dW, Z, lcut):
        a, b = self._g(), self._g()
        self.lib.orc_torpol_to_dphspat(self.h, _p(self._c(dW)), _p(self._c(Z)), _p(a), _p(b), c_int(lcut))
        return a, b

    def pol_to_curlr_spat(self, Q, lcut):
        a = self._g()
        self.lib.orc_pol_to_curlr_spat(self.h, _p(self._c(Q)), _p(a), c_int(lcut))
        return a

    def torpol_to_curl_spat(self, or2, B, ddB, J, dJ, lcut):
        a, b, c = self._g(), self._g(), self._g()
        self.lib.orc_torpol_to_curl_spat(self.h, c_double(or2), _p(self._c(B)), _p(self._c(ddB)), _p(self._c(J)),
                                         _p(self._c(dJ)), _p(a), _p(b), _p(c), c_int(lcut))
        return a, b, c

    def scal_to_SH(self, f, lcut):
        o = self._s()
        self.lib.orc_scal_to_SH(self.h, _p(self._r(f)), _p(o), c_int(lcut))
        return o

    def spat_to_qst(self, f, g, h, lcut):
        q, s, t = self._s(), self._s(), self._s()
        self.lib.orc_spat_to_qst(self.h, _p(self._r(f)), _p(self._r(g)), _p(self._r(h)), _p(q), _p(s), _p(t),
                                 c_int(lcut))
        return q, s, t

    def spat_to_sphertor(self, f, g, lcut):
        s, t = self._s(), self._s()
        self.lib.orc_spat_to_sphertor(self.h, _p(self._r(f)), _p(self._r(g)), _p(s), _p(t), c_int(lcut))
        return s, t

    def torpol_to_spat_IC(self, r, r_ICB, W, dW, Z):
        a, b, c = self._g(), self._g(), self._g()
        self.lib.orc_torpol_to_spat_IC(self.h, c_double(r), c_double(r_ICB), _p(self._c(W)), _p(self._c(dW)),
                                       _p(self._c(Z)), _p(a), _p(b), _p(c))
        return a, b, c

    def torpol_to_curl_spat_IC(self, r, r_ICB, dB, ddB, J, dJ):
        a, b, c = self._g(), self._g(), self._g()
        self.lib.orc_torpol_to_curl_spat_IC(self.h, c_double(r), c_double(r_ICB), _p(self._c(dB)), _p(self._c(ddB)),
                                            _p(self._c(J)), _p(self._c(dJ)), _p(a), _p(b), _p(c))
        return a, b, c

    def axi_to_spat(self, fl_ax):
        f = np.zeros(self.n_theta)
        self.lib.orc_axi_to_spat(self.h, _p(self._c(fl_ax)), _p(f))
        return f

    def toraxi_to_spat(self, fl_ax, lcut):
        a, b = np.zeros(self.n_theta), np.zeros(self.n_theta)
        self.lib.orc_toraxi_to_spat(self.h, _p(self._c(fl_ax)), _p(a), _p(b), c_int(lcut))
        return a, b

    # fft.f90 -----------------------------------------------------------------------------------
    def ifft_many(self, f):
        g = self._g()
        self.lib.orc_ifft_many(self.h, _p(self._c(f)), _p(g))
        return g

    def fft_many(self, g):
        f = np.zeros((self.n_phi // 2 + 1, self.n_theta), dtype=np.complex128)
        self.lib.orc_fft_many(self.h, _p(self._r(g)), _p(f))
        return f

    # radial loop -------------------------------------------------------------------------------
    def radial_loop(self, params, radial, fields, time=0.0):
        """rIter.f90:94-464 for the levels in `radial` (dict of per-level arrays incl. 'nR','l_R').
        fields: dict name -> complex128 [n_r, lm_max].  Returns dict of outputs."""
        n_r = len(radial["nR"])
        keep = []
        rad = _Radial()
        for nm in ["nR", "l_R"]:
            a = np.ascontiguousarray(radial[nm], dtype=np.int32)
            keep.append(a)
            setattr(rad, nm, _p(a))
        for nm in _RAD_NAMES:
            key = "lambda" if nm == "lambda_" else nm
            a = np.ascontiguousarray(radial.get(key, np.ones(n_r)), dtype=np.float64)
            keep.append(a)
            setattr(rad, nm, _p(a))
        fin = _FieldsIn()
        for nm in _IN_NAMES:
            if nm in fields and fields[nm] is not None:
                a = self._c(fields[nm])
                assert a.shape == (n_r, self.lm_max)
                keep.append(a)
                setattr(fin, nm, _p(a))
        out = {}
        fout = _FieldsOut()
        for nm in _OUT_NAMES:
            out[nm] = np.zeros((n_r, self.lm_max), dtype=np.complex128)
            setattr(fout, nm, _p(out[nm]))
        out["dtrkc"] = np.zeros(n_r)
        out["dthkc"] = np.zeros(n_r)
        fout.dtrkc = _p(out["dtrkc"])
        fout.dthkc = _p(out["dthkc"])
        tq = np.zeros(2)
        fout.lorentz_torque_ic = tq.ctypes.data
        fout.lorentz_torque_ma = tq.ctypes.data + 8
        for nm in ("br_vt_lm_cmb", "br_vp_lm_cmb", "br_vt_lm_icb", "br_vp_lm_icb"):  # get_br_v_bcs, rIter.f90:267-277
            out[nm] = np.zeros(self.lm_max, dtype=np.complex128)
            setattr(fout, nm, _p(out[nm]))
        if getattr(params, "l_phase_field", 0):
            out["dphidt"] = np.zeros((n_r, self.lm_max), dtype=np.complex128)
            fout.dphidt = _p(out["dphidt"])
        self.lib.orc_radial_loop(self.h, C.byref(params), C.byref(rad), c_int(n_r), C.byref(fin), C.byref(fout),
                                 c_double(time))
        out["lorentz_torque_ic"], out["lorentz_torque_ma"] = float(tq[0]), float(tq[1])
        return out

    def radial_diagnostics(self, params, radial, fields, mask, ktops=1, kbots=1):
        """rIter.f90:303-373 (get_helicity, get_hemi, get_visc_heat, get_perpPar, get_fluxes, get_nlBLayers) for the levels in
        `radial`: float64 [n_r, 40], slots as MAGIC_DG_* of include/magic_sht.h."""
        n_r = len(radial["nR"])
        keep = []
        rad = _Radial()
        for nm in ["nR", "l_R"]:
            a = np.ascontiguousarray(radial[nm], dtype=np.int32)
            keep.append(a)
            setattr(rad, nm, _p(a))
        for nm in _RAD_NAMES:
            key = "lambda" if nm == "lambda_" else nm
            a = np.ascontiguousarray(radial.get(key, np.ones(n_r)), dtype=np.float64)
            keep.append(a)
            setattr(rad, nm, _p(a))
        fin = _FieldsIn()
        for nm in _IN_NAMES:
            if nm in fields and fields[nm] is not None:
                a = self._c(fields[nm])
                assert a.shape == (n_r, self.lm_max)
                keep.append(a)
                setattr(fin, nm, _p(a))
        out = np.zeros((n_r, 40))
        self.lib.orc_radial_diagnostics(self.h, C.byref(params), C.byref(rad), c_int(n_r), C.byref(fin), c_int(mask), c_int(ktops),
                                        c_int(kbots), _p(out))
        return out

    def radial_dtB(self, params, radial, fields):
        """get_dtBLM (dtB.f90:144-223) for the levels in `radial`: complex128 [11, n_r, lm_max] (BtVrLM, BpVrLM, BrVtLM, BrVpLM,
        BtVpLM, BpVtLM, BpVtBtVpCotLM, BpVtBtVpSn2LM, BrVZLM, BtVZLM, BtVZsn2LM)."""
        n_r = len(radial["nR"])
        keep = []
        rad = _Radial()
        for nm in ["nR", "l_R"]:
            a = np.ascontiguousarray(radial[nm], dtype=np.int32)
            keep.append(a)
            setattr(rad, nm, _p(a))
        for nm in _RAD_NAMES:
            key = "lambda" if nm == "lambda_" else nm
            a = np.ascontiguousarray(radial.get(key, np.ones(n_r)), dtype=np.float64)
            keep.append(a)
            setattr(rad, nm, _p(a))
        fin = _FieldsIn()
        for nm in _IN_NAMES:
            if nm in fields and fields[nm] is not None:
                a = self._c(fields[nm])
                keep.append(a)
                setattr(fin, nm, _p(a))
        out = np.zeros((11, n_r, self.lm_max), dtype=np.complex128)
        self.lib.orc_radial_dtB(self.h, C.byref(params), C.byref(rad), c_int(n_r), C.byref(fin), _p(out))
        return out

    def radial_TO(self, params, radial, fields, mode, dtLast=1.0, last=None):
        """rIter.f90:395-404.  mode 0: getTOnext's grid part (TO.f90:330-343), returns last = float64 [n_r, 3, n_phi, n_theta]
        (BsLast, BpLast, BzLast); mode 1: getTO (TO.f90:141-307) with that `last`, returns float64 [n_r, 15, n_theta] (colatitudes
        north -> south; V2AS, VAS, dzCorAS, dzRstrAS, dzAstrAS, dzLFAS, Bs2AS, BspAS, BpzAS, BszAS, BspdAS, BpsdAS, BzpdAS, BpzdAS,
        dzPenAS)."""
        n_r = len(radial["nR"])
        keep = []
        rad = _Radial()
        for nm in ["nR", "l_R"]:
            a = np.ascontiguousarray(radial[nm], dtype=np.int32)
            keep.append(a)
            setattr(rad, nm, _p(a))
        for nm in _RAD_NAMES:
            key = "lambda" if nm == "lambda_" else nm
            a = np.ascontiguousarray(radial.get(key, np.ones(n_r)), dtype=np.float64)
            keep.append(a)
            setattr(rad, nm, _p(a))
        fin = _FieldsIn()
        for nm in _IN_NAMES:
            if nm in fields and fields[nm] is not None:
                a = self._c(fields[nm])
                keep.append(a)
                setattr(fin, nm, _p(a))
        if last is None:
            last = np.zeros((n_r, 3, self.n_phi, self.n_theta))
        last = np.ascontiguousarray(last, dtype=np.float64)
        assert last.shape == (n_r, 3, self.n_phi, self.n_theta)
        out = np.zeros((n_r, 15, self.n_theta))
        self.lib.orc_radial_TO(self.h, C.byref(params), C.byref(rad), c_int(n_r), C.byref(fin), c_int(mode), c_double(dtLast), _p(last),
                               _p(out))
        return last if mode == 0 else out

    def radial_RMS(self, params, radial, fields, old, dt, time=0.0):
        """The r.m.s. force balance inside the radial loop on lRmsCalc steps (rIter.f90:215-252, 710; RMS.f90:469-610): complex128
        [14, n_r, lm_max] (AdvrLM, LFrLM, dtVrLM, dpkindrLM, Advt2LM, Advp2LM, LFt2LM, LFp2LM, CFt2LM, CFp2LM, PFt2LM, PFp2LM, dtVtLM,
        dtVpLM).  old: dict with the w, dw, z of the previous stage-1 call."""
        n_r = len(radial["nR"])
        keep = []
        rad = _Radial()
        for nm in ["nR", "l_R"]:
            a = np.ascontiguousarray(radial[nm], dtype=np.int32)
            keep.append(a)
            setattr(rad, nm, _p(a))
        for nm in _RAD_NAMES:
            key = "lambda" if nm == "lambda_" else nm
            a = np.ascontiguousarray(radial.get(key, np.ones(n_r)), dtype=np.float64)
            keep.append(a)
            setattr(rad, nm, _p(a))
        fin = _FieldsIn()
        for nm in _IN_NAMES:
            if nm in fields and fields[nm] is not None:
                a = self._c(fields[nm])
                keep.append(a)
                setattr(fin, nm, _p(a))
        o = [self._c(old[k]) for k in ("w", "dw", "z")]
        out = np.zeros((14, n_r, self.lm_max), dtype=np.complex128)
        self.lib.orc_radial_RMS(self.h, C.byref(params), C.byref(rad), c_int(n_r), C.byref(fin), _p(o[0]), _p(o[1]), _p(o[2]), c_double(dt),
                                c_double(time), _p(out))
        return out

    def get_nl_mhd(self, params, nR, nBc, or2, or4, orho1, grids_in):
        """get_nl.f90:213-441 on 13 caller grids -> 12 product grids."""
        ins = [self._r(g) for g in grids_in]
        outs = [self._g() for _ in range(12)]
        pin = (c_void_p * 13)(*[_p(a) for a in ins])
        pout = (c_void_p * 12)(*[_p(a) for a in outs])
        self.lib.orc_get_nl_mhd(self.h, C.byref(params), c_int(nR), c_int(nBc), c_double(or2), c_double(or4),
                                c_double(orho1), pin, pout)
        return outs

    # mpi_transpose.f90 (alltoallv flavour, emulated in-process) -----------------------------------
    def transp_lm2r(self, n_procs, n_r_max, arr_LM):
        """arr_LM[p]: complex128 [n_fields, n_r_max, nlm_loc(p)] -> list arr_R[q] [n_fields, nR_loc(q), lm_max]."""
        n_fields = arr_LM[0].shape[0]
        rs, re = get_blocks(n_r_max, n_procs)
        ins = [self._c(a) for a in arr_LM]
        outs = [np.zeros((n_fields, re[q] - rs[q] + 1, self.lm_max), dtype=np.complex128) for q in range(n_procs)]
        pin = (c_void_p * n_procs)(*[_p(a) for a in ins])
        pout = (c_void_p * n_procs)(*[_p(a) for a in outs])
        self.lib.orc_transp_lm2r(self.h, c_int(n_procs), c_int(n_r_max), c_int(n_fields), pin, pout)
        return outs

    def transp_r2lm(self, n_procs, n_r_max, arr_R):
        n_fields = arr_R[0].shape[0]
        _, ls, le = self.lo_map(n_procs)
        ins = [self._c(a) for a in arr_R]
        outs = [np.zeros((n_fields, n_r_max, le[p] - ls[p] + 1), dtype=np.complex128) for p in range(n_procs)]
        pin = (c_void_p * n_procs)(*[_p(a) for a in ins])
        pout = (c_void_p * n_procs)(*[_p(a) for a in outs])
        self.lib.orc_transp_r2lm(self.h, c_int(n_procs), c_int(n_r_max), c_int(n_fields), pin, pout)
        return outs
