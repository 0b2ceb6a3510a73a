"""ctypes binding of libmagic_b200.so (the C ABI declared in include/magic_sht.h)."""
import ctypes as C
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_longlong, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MAGIC_B200_LIB", os.path.join(_HERE, "libmagic_b200.so"))  # env override: experimental builds
_lib = None


class MagicError(RuntimeError):
    """Raised when a C-ABI call returns non-zero (the Fortran shim would call abortRun, useful.f90:271)."""


# every symbol include/magic_sht.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "magic_last_error", "magic_device_count", "magic_sht_create", "magic_sht_destroy", "magic_sht_get_grid", "magic_sht_stream", "magic_sht_launch_count",
    "magic_scal_to_spat", "magic_scal_to_grad_spat", "magic_pol_to_grad_spat", "magic_torpol_to_spat",
    "magic_sphtor_to_spat", "magic_torpol_to_curl_spat_IC", "magic_torpol_to_spat_IC", "magic_torpol_to_dphspat",
    "magic_pol_to_curlr_spat", "magic_torpol_to_curl_spat", "magic_scal_to_SH", "magic_spat_to_qst",
    "magic_spat_to_sphertor", "magic_axi_to_spat", "magic_toraxi_to_spat",
    "magic_rloop_create", "magic_rloop_destroy", "magic_level_chunks", "magic_rloop_run", "magic_rloop_run_dev", "magic_rloop_sync",
    "magic_rloop_set_rotation", "magic_rloop_get_torques", "magic_rloop_get_br_v_bcs", "magic_rloop_diagnostics", "magic_rloop_diagnostics_dev", "magic_rloop_graph_fields", "magic_rloop_dtb", "magic_rloop_dtb_dev", "magic_rloop_to_next", "magic_rloop_to_next_dev", "magic_rloop_to", "magic_rloop_to_dev", "magic_rloop_rms_keep", "magic_rloop_rms_keep_dev", "magic_rloop_rms", "magic_rloop_rms_dev", "magic_rloop_launch_count", "magic_rloop_last_timing", "magic_rloop_last_exposed", "magic_rloop_legendre_flops", "magic_rloop_legendre_units", "magic_rloop_pin_host", "magic_rloop_unpin_host", "magic_rloop_level_chunk",
    "magic_transp_unique_id", "magic_transp_create", "magic_transp_destroy", "magic_transp_extents",
    "magic_transp_create_part", "magic_transp_set_stream", "magic_transp_info", "magic_rloop_run_lm_dev", "magic_rloop_run_lm", "magic_rloop_set_radial_matrices", "magic_rloop_set_lm_radial", "magic_rloop_lm_options",
    "magic_transp_lm2r_dev", "magic_transp_r2lm_dev", "magic_transp_lm2r_dev_n", "magic_transp_r2lm_dev_n", "magic_transp_lm2r", "magic_transp_r2lm",
    "magic_transp_pack_lm2r_dev", "magic_transp_unpack_lm2r_dev", "magic_transp_pack_r2lm_dev",
    "magic_transp_unpack_r2lm_dev", "magic_transp_counts",
    "magic_get_blocks", "magic_lo_map",
    "magic_dev_malloc", "magic_dev_free", "magic_dev_upload", "magic_dev_download",
]


def load_library():
    """Loads the CUDA library; raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MagicError(f"{LIB_PATH} is missing: run `python -m magic_b200.build` (magic_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.magic_last_error.restype = c_char_p
    lib.magic_rloop_launch_count.restype = c_longlong
    lib.magic_rloop_launch_count.argtypes = [c_void_p]
    lib.magic_rloop_legendre_flops.restype = c_double
    lib.magic_rloop_legendre_flops.argtypes = [c_void_p]
    lib.magic_rloop_pin_host.argtypes = [c_void_p, c_void_p, c_size_t]
    lib.magic_rloop_unpin_host.argtypes = [c_void_p, c_void_p]
    lib.magic_rloop_level_chunk.argtypes = [c_void_p]
    lib.magic_sht_stream.restype = c_void_p
    lib.magic_sht_stream.argtypes = [c_void_p]
    lib.magic_sht_launch_count.restype = c_longlong
    lib.magic_sht_launch_count.argtypes = [c_void_p]
    lib.magic_dev_malloc.argtypes = [c_void_p, c_size_t, POINTER(c_void_p)]
    lib.magic_dev_free.argtypes = [c_void_p, c_void_p]
    lib.magic_dev_upload.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t]
    lib.magic_dev_download.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t]
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise MagicError(load_library().magic_last_error().decode())


def ptr(a):
    """numpy array or int device pointer or None -> c_void_p."""
    if a is None:
        return c_void_p(None)
    if isinstance(a, int):
        return c_void_p(a)
    return a.ctypes.data_as(c_void_p)
