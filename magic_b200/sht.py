"""Host-side mirror of MagIC's `module sht` (sht_native.f90:16-20 == shtns.f90:22-26).

Same procedure names, argument order and meaning as the Fortran subroutines; arrays are numpy:
  spectra  complex128 [lm_max]               (st_map order, blocking.f90:309-317)
  grids    float64    [n_phi_max, nlat_padded] C order == Fortran f(nlat_padded, n_phi_max), theta rows
           N/S interleaved (initialize_sht returns l_scrambled_theta=.true., sht_native.f90:29).
Every call goes through the C ABI (include/magic_sht.h) with host buffers, i.e. the path a Fortran shim takes.
"""
from ctypes import POINTER, byref, c_double, c_int, c_void_p

import numpy as np

from .lib import MagicError, check, load_library, ptr


def grid_sizes(l_max=0, n_phi_tot=0, minc=1, nalias=20):
    """truncation.f90:55-105: grid sizes from l_max (or from n_phi_tot when l_max == 0)."""
    def prime_decomposition(nlon):  # truncation.f90:162-192
        best = None
        for i in range(13):
            for j in range(7):
                for k in range(7):
                    res = 2 ** i * 3 ** j * 5 ** k
                    d = res - nlon
                    if 0 <= d < 100 and (best is None or d < best[0]):
                        best = (d, res)
        return best[1]

    if l_max == 0:
        n_phi_max = n_phi_tot // minc
        n_theta_max = n_phi_tot // 2
        l_max = (nalias * n_theta_max) // 30
    else:
        n_theta_max = (30 * l_max) // nalias
        n_phi_tot = prime_decomposition(2 * n_theta_max)
        n_phi_max = n_phi_tot // minc
        n_theta_max = n_phi_tot // 2
    m_max = min((l_max // minc) * minc, l_max)
    n_m_max = m_max // minc + 1
    lm_max = sum(l_max - m + 1 for m in range(0, m_max + 1, minc))
    return dict(l_max=l_max, m_max=m_max, n_theta_max=n_theta_max, n_phi_max=n_phi_max, n_m_max=n_m_max,
                lm_max=lm_max, n_phi_tot=n_phi_tot)


class Sht:
    """initialize_sht ... finalize_sht (sht_native.f90:24-38)."""

    def __init__(self, l_max, m_max=None, minc=1, n_theta_max=None, n_phi_max=None, nlat_padded=None, device_id=0,
                 nalias=20):
        self.lib = load_library()
        if n_theta_max is None:
            gs = grid_sizes(l_max=l_max, minc=minc, nalias=nalias)
            n_theta_max, n_phi_max = gs["n_theta_max"], gs["n_phi_max"]
        if m_max is None:
            m_max = (l_max // minc) * minc
        if nlat_padded is None:
            nlat_padded = n_theta_max
        self.l_max, self.m_max, self.minc = l_max, m_max, minc
        self.n_theta_max, self.n_phi_max, self.nlat_padded = n_theta_max, n_phi_max, nlat_padded
        self.n_m_max = m_max // minc + 1
        self.lm_max = sum(l_max - m + 1 for m in range(0, m_max + 1, minc))
        self.lm2l = np.concatenate([np.arange(m, l_max + 1) for m in range(0, m_max + 1, minc)]).astype(np.int32)
        self.lm2m = np.concatenate([np.full(l_max - m + 1, m) for m in range(0, m_max + 1, minc)]).astype(np.int32)
        self._h = c_void_p()
        scr = c_int(0)
        check(self.lib.magic_sht_create(c_int(l_max), c_int(m_max), c_int(minc), c_int(n_theta_max), c_int(n_phi_max),
                                        c_int(nlat_padded), c_int(device_id), byref(scr), byref(self._h)))
        self.l_scrambled_theta = bool(scr.value)

    # -- lifetime ------------------------------------------------------------------------------------
    def finalize_sht(self):
        if self._h:
            self.lib.magic_sht_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.finalize_sht()
        except Exception:
            pass

    @property
    def handle(self):
        if not self._h:
            raise MagicError("Sht handle already finalized")
        return self._h

    @property
    def stream(self):
        """cudaStream_t (int) all work of this handle is issued on."""
        return int(self.lib.magic_sht_stream(self.handle) or 0)

    def launch_count(self):
        return int(self.lib.magic_sht_launch_count(self.handle))

    def get_grid(self):
        th = np.zeros(self.n_theta_max)
        g = np.zeros(self.n_theta_max)
        check(self.lib.magic_sht_get_grid(self.handle, ptr(th), ptr(g)))
        return th, g

    # -- helpers -------------------------------------------------------------------------------------
    def _spec(self, a):
        a = np.ascontiguousarray(a, dtype=np.complex128)
        if a.shape != (self.lm_max,):
            raise ValueError(f"spectral array must have shape ({self.lm_max},), got {a.shape}")
        return a

    def _grid(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.shape != (self.n_phi_max, self.nlat_padded):
            raise ValueError(f"grid array must have shape ({self.n_phi_max}, {self.nlat_padded}), got {a.shape}")
        return a

    def _g(self):
        return np.zeros((self.n_phi_max, self.nlat_padded))

    def _s(self):
        return np.zeros(self.lm_max, dtype=np.complex128)

    # -- the 17 procedures ------------------------------------------------------------------------------
    def scal_to_spat(self, Slm, lcut):
        f = self._g()
        check(self.lib.magic_scal_to_spat(self.handle, ptr(self._spec(Slm)), ptr(f), c_int(lcut)))
        return f

    def scal_to_grad_spat(self, Slm, lcut):
        a, b = self._g(), self._g()
        check(self.lib.magic_scal_to_grad_spat(self.handle, ptr(self._spec(Slm)), ptr(a), ptr(b), c_int(lcut)))
        return a, b

    def pol_to_grad_spat(self, Slm, lcut):
        a, b = self._g(), self._g()
        check(self.lib.magic_pol_to_grad_spat(self.handle, ptr(self._spec(Slm)), ptr(a), ptr(b), c_int(lcut)))
        return a, b

    def torpol_to_spat(self, Wlm, dWlm, Zlm, lcut):
        a, b, c = self._g(), self._g(), self._g()
        check(self.lib.magic_torpol_to_spat(self.handle, ptr(self._spec(Wlm)), ptr(self._spec(dWlm)), ptr(self._spec(Zlm)),
                                            ptr(a), ptr(b), ptr(c), c_int(lcut)))
        return a, b, c

    def sphtor_to_spat(self, dWlm, Zlm, lcut):
        a, b = self._g(), self._g()
        check(self.lib.magic_sphtor_to_spat(self.handle, ptr(self._spec(dWlm)), ptr(self._spec(Zlm)), ptr(a), ptr(b),
                                            c_int(lcut)))
        return a, b

    def torpol_to_dphspat(self, dWlm, Zlm, lcut):
        a, b = self._g(), self._g()
        check(self.lib.magic_torpol_to_dphspat(self.handle, ptr(self._spec(dWlm)), ptr(self._spec(Zlm)), ptr(a), ptr(b),
                                               c_int(lcut)))
        return a, b

    def pol_to_curlr_spat(self, Qlm, lcut):
        a = self._g()
        check(self.lib.magic_pol_to_curlr_spat(self.handle, ptr(self._spec(Qlm)), ptr(a), c_int(lcut)))
        return a

    def torpol_to_curl_spat(self, or2, Blm, ddBlm, Jlm, dJlm, lcut):
        a, b, c = self._g(), self._g(), self._g()
        check(self.lib.magic_torpol_to_curl_spat(self.handle, c_double(or2), ptr(self._spec(Blm)), ptr(self._spec(ddBlm)),
                                                 ptr(self._spec(Jlm)), ptr(self._spec(dJlm)), ptr(a), ptr(b), ptr(c),
                                                 c_int(lcut)))
        return a, b, c

    def torpol_to_spat_IC(self, r, r_ICB, Wlm, dWlm, Zlm):
        a, b, c = self._g(), self._g(), self._g()
        check(self.lib.magic_torpol_to_spat_IC(self.handle, c_double(r), c_double(r_ICB), ptr(self._spec(Wlm)),
                                               ptr(self._spec(dWlm)), ptr(self._spec(Zlm)), ptr(a), ptr(b), ptr(c)))
        return a, b, c

    def torpol_to_curl_spat_IC(self, r, r_ICB, dBlm, ddBlm, Jlm, dJlm):
        a, b, c = self._g(), self._g(), self._g()
        check(self.lib.magic_torpol_to_curl_spat_IC(self.handle, c_double(r), c_double(r_ICB), ptr(self._spec(dBlm)),
                                                    ptr(self._spec(ddBlm)), ptr(self._spec(Jlm)), ptr(self._spec(dJlm)),
                                                    ptr(a), ptr(b), ptr(c)))
        return a, b, c

    def scal_to_SH(self, f, lcut):
        o = self._s()
        check(self.lib.magic_scal_to_SH(self.handle, ptr(self._grid(f)), ptr(o), c_int(lcut)))
        return o

    def spat_to_qst(self, f, g, h, lcut):
        q, s, t = self._s(), self._s(), self._s()
        check(self.lib.magic_spat_to_qst(self.handle, ptr(self._grid(f)), ptr(self._grid(g)), ptr(self._grid(h)), ptr(q),
                                         ptr(s), ptr(t), c_int(lcut)))
        return q, s, t

    def spat_to_sphertor(self, f, g, lcut):
        s, t = self._s(), self._s()
        check(self.lib.magic_spat_to_sphertor(self.handle, ptr(self._grid(f)), ptr(self._grid(g)), ptr(s), ptr(t),
                                              c_int(lcut)))
        return s, t

    def axi_to_spat(self, fl_ax):
        a = np.ascontiguousarray(fl_ax, dtype=np.complex128)
        f = np.zeros(self.n_theta_max)
        check(self.lib.magic_axi_to_spat(self.handle, ptr(a), ptr(f)))
        return f

    def toraxi_to_spat(self, fl_ax, lcut):
        a = np.ascontiguousarray(fl_ax, dtype=np.complex128)
        ft, fp = np.zeros(self.n_theta_max), np.zeros(self.n_theta_max)
        check(self.lib.magic_toraxi_to_spat(self.handle, ptr(a), ptr(ft), ptr(fp), c_int(lcut)))
        return ft, fp
