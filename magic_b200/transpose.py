"""Host-side mirror of `type_mpitransp` (mpi_transpose.f90:18-54): create_comm / transp_lm2r / transp_r2lm."""
from ctypes import byref, c_char, c_int, c_longlong, c_void_p, create_string_buffer

import numpy as np

from .lib import MagicError, check, load_library, ptr


def get_blocks(n_points, n_procs):
    """getBlocks (parallel.f90:75-92): 1-based inclusive (start, stop) per rank; remainder on the last ranks."""
    n_loc = n_points // n_procs
    rem = n_points - n_loc * n_procs
    start, stop = [], []
    for p in range(n_procs):
        s = n_loc * p + max(p + rem - n_procs, 0) + 1
        e = n_loc * (p + 1) + max(p + rem + 1 - n_procs, 0)
        if p != 0:
            s = stop[-1] + 1
        start.append(s)
        stop.append(e)
    return np.array(start), np.array(stop)


def lo_map(l_max, m_max, minc, n_procs):
    """lo_map of blocking.f90:339-544 from the library (host only): (lo2st, lm_start, lm_stop)."""
    lib = load_library()
    lm_max = sum(l_max - m + 1 for m in range(0, m_max + 1, minc))
    lo2st = np.zeros(lm_max, dtype=np.int32)
    s = np.zeros(n_procs, dtype=np.int32)
    e = np.zeros(n_procs, dtype=np.int32)
    check(lib.magic_lo_map(c_int(l_max), c_int(m_max), c_int(minc), c_int(n_procs), ptr(lo2st), ptr(s), ptr(e)))
    return lo2st, s, e


def unique_id():
    """NCCL bootstrap id (128 bytes) -- create on rank 0 and broadcast."""
    lib = load_library()
    buf = create_string_buffer(128)
    check(lib.magic_transp_unique_id(buf))
    return buf.raw


class Transposer:
    """One container transposer (e.g. lo2r_flow / r2lo_flow of communications.f90:66-69)."""

    def __init__(self, sht, n_r_max, n_fields, rank=0, n_procs=1, nccl_id=None):
        self.lib = load_library()
        self.sht, self.rank, self.n_procs, self.n_r_max, self.n_fields = sht, rank, n_procs, n_r_max, n_fields
        self._h = c_void_p()
        idbuf = create_string_buffer(nccl_id, 128) if nccl_id is not None else None
        check(self.lib.magic_transp_create(sht.handle, idbuf, c_int(rank), c_int(n_procs), c_int(n_r_max), c_int(n_fields),
                                           byref(self._h)))
        a, b, c, d = c_int(), c_int(), c_int(), c_int()
        check(self.lib.magic_transp_extents(self._h, byref(a), byref(b), byref(c), byref(d)))
        self.llm, self.ulm, self.nRstart, self.nRstop = a.value, b.value, c.value, d.value
        self.nlm_loc = self.ulm - self.llm + 1
        self.nr_loc = self.nRstop - self.nRstart + 1

    def part(self, lev_off, lev_cnt):
        """magic_transp_create_part: a transposer that moves, for every rank q, only the levels
        [lev_off[q], lev_off[q]+lev_cnt[q]) of q's slab (one level chunk).  Array shapes stay those of the parent."""
        child = object.__new__(Transposer)
        child.lib, child.sht, child.rank, child.n_procs = self.lib, self.sht, self.rank, self.n_procs
        child.n_r_max, child.n_fields = self.n_r_max, self.n_fields
        child.llm, child.ulm, child.nlm_loc, child.nr_loc = self.llm, self.ulm, self.nlm_loc, self.nr_loc
        off = (c_int * self.n_procs)(*[int(x) for x in lev_off])
        cnt = (c_int * self.n_procs)(*[int(x) for x in lev_cnt])
        child._h = c_void_p()
        check(self.lib.magic_transp_create_part(self._h, off, cnt, byref(child._h)))
        child.nRstart, child.nRstop = self.nRstart + int(lev_off[self.rank]), self.nRstart + int(lev_off[self.rank]) + int(lev_cnt[self.rank]) - 1
        child._parent = self  # keep the parent alive
        return child

    def destroy_comm(self):
        if self._h:
            self.lib.magic_transp_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.destroy_comm()
        except Exception:
            pass

    def counts(self, direction):
        n = self.n_procs
        arrs = [(c_longlong * n)() for _ in range(4)]
        check(self.lib.magic_transp_counts(self._h, c_int(direction), *arrs))
        return [np.array(a) for a in arrs]

    # host-buffer calls: arr_LMloc [n_fields, n_r_max, nlm_loc], arr_Rloc [n_fields, nr_loc, lm_max]
    def transp_lm2r(self, arr_LMloc):
        a = np.ascontiguousarray(arr_LMloc, dtype=np.complex128)
        assert a.shape == (self.n_fields, self.n_r_max, self.nlm_loc)
        out = np.zeros((self.n_fields, self.nr_loc, self.sht.lm_max), dtype=np.complex128)
        check(self.lib.magic_transp_lm2r(self._h, ptr(a), ptr(out)))
        return out

    def transp_r2lm(self, arr_Rloc):
        a = np.ascontiguousarray(arr_Rloc, dtype=np.complex128)
        assert a.shape == (self.n_fields, self.nr_loc, self.sht.lm_max)
        out = np.zeros((self.n_fields, self.n_r_max, self.nlm_loc), dtype=np.complex128)
        check(self.lib.magic_transp_r2lm(self._h, ptr(a), ptr(out)))
        return out

    # device-pointer calls (ints)
    def transp_lm2r_dev(self, arr_LMloc_dev, arr_Rloc_dev):
        check(self.lib.magic_transp_lm2r_dev(self._h, c_void_p(arr_LMloc_dev), c_void_p(arr_Rloc_dev)))

    def transp_r2lm_dev(self, arr_Rloc_dev, arr_LMloc_dev):
        check(self.lib.magic_transp_r2lm_dev(self._h, c_void_p(arr_Rloc_dev), c_void_p(arr_LMloc_dev)))

    def transp_lm2r_dev_n(self, n_fields, arr_LMloc_dev, arr_Rloc_dev):
        check(self.lib.magic_transp_lm2r_dev_n(self._h, c_int(n_fields), c_void_p(arr_LMloc_dev), c_void_p(arr_Rloc_dev)))

    def transp_r2lm_dev_n(self, n_fields, arr_Rloc_dev, arr_LMloc_dev):
        check(self.lib.magic_transp_r2lm_dev_n(self._h, c_int(n_fields), c_void_p(arr_Rloc_dev), c_void_p(arr_LMloc_dev)))

    def pack_lm2r_dev(self, arr_dev, buf_dev):
        check(self.lib.magic_transp_pack_lm2r_dev(self._h, c_void_p(arr_dev), c_void_p(buf_dev)))

    def unpack_lm2r_dev(self, buf_dev, arr_dev):
        check(self.lib.magic_transp_unpack_lm2r_dev(self._h, c_void_p(buf_dev), c_void_p(arr_dev)))

    def pack_r2lm_dev(self, arr_dev, buf_dev):
        check(self.lib.magic_transp_pack_r2lm_dev(self._h, c_void_p(arr_dev), c_void_p(buf_dev)))

    def unpack_r2lm_dev(self, buf_dev, arr_dev):
        check(self.lib.magic_transp_unpack_r2lm_dev(self._h, c_void_p(buf_dev), c_void_p(arr_dev)))
