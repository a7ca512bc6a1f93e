"""ctypes binding of the C ABI in include/phylocsf_b200.h. Loading fails loudly when the CUDA
library has not been built; there is no Python/CPU fallback for any compute entry point."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PCSF_LIB selects another build of the same library (kernel A/B experiments); never a different implementation
LIB_PATH = os.environ.get("PCSF_LIB") or os.path.join(_HERE, "libphylocsf_b200.so")

# every symbol include/phylocsf_b200.h declares
SYMBOLS = [
    "pcsf_version", "pcsf_device_count", "pcsf_create", "pcsf_destroy", "pcsf_last_error", "pcsf_stream_set", "pcsf_option_set",
    "pcsf_tree_set", "pcsf_model_set", "pcsf_pt_build", "pcsf_pt_get", "pcsf_batch_upload",
    "pcsf_batch_upload_alignments", "pcsf_batch_upload_alignments_parts", "pcsf_batch_nregions", "pcsf_batch_ncols", "pcsf_batch_codes_get", "pcsf_lpr_all", "pcsf_score_alignments", "pcsf_lpr",
    "pcsf_models_set", "pcsf_omega_models_set", "pcsf_model_get", "pcsf_pt_build_pairs", "pcsf_lpr_pairs", "pcsf_column_terms", "pcsf_maximize_lpr", "pcsf_maximize_lpr_multi", "pcsf_last_ms", "pcsf_launch_count", "pcsf_table_level", "pcsf_last_launch_info", "pcsf_tree_n_leaves", "pcsf_total_ms", "pcsf_posteriors", "pcsf_counter", "pcsf_omega_cache_reset", "pcsf_omega_models_set_cached",
]

PCSF_OK = 0
ST_NEG_T, ST_NEG_ENTRY, ST_ROWSUM, ST_DIAG_ASSERT, ST_NOT_FINITE, ST_BRACKET, ST_RANDOM_INIT = 1, 2, 4, 8, 16, 32, 64

_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            "%s not found: build it with `python -m phylocsf_b200.build` (nvcc, sm_100a). "
            "phylocsf_b200 has no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
    L.pcsf_version.restype = ctypes.c_char_p
    L.pcsf_device_count.restype = ctypes.c_int
    L.pcsf_create.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
    L.pcsf_destroy.argtypes = [vp]
    L.pcsf_destroy.restype = None
    L.pcsf_last_error.argtypes = [vp]
    L.pcsf_last_error.restype = ctypes.c_char_p
    L.pcsf_stream_set.argtypes = [vp, vp]
    L.pcsf_option_set.argtypes = [vp, ctypes.c_int, i64]
    L.pcsf_tree_set.argtypes = [vp, ctypes.c_int, vp, vp]
    L.pcsf_model_set.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp]
    L.pcsf_pt_build.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp]
    L.pcsf_pt_get.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]
    L.pcsf_batch_upload.argtypes = [vp, i64, vp, vp]
    L.pcsf_batch_upload_alignments.argtypes = [vp, i64, vp, vp, vp, ctypes.c_int]
    L.pcsf_batch_upload_alignments_parts.argtypes = [vp, i64, vp, vp, i64, vp, vp, ctypes.c_int]
    L.pcsf_batch_codes_get.argtypes = [vp, vp]
    L.pcsf_batch_nregions.argtypes = [vp]
    L.pcsf_batch_nregions.restype = i64
    L.pcsf_batch_ncols.argtypes = [vp]
    L.pcsf_batch_ncols.restype = i64
    L.pcsf_lpr_all.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, vp]
    L.pcsf_score_alignments.argtypes = [vp, i64, vp, vp, vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp]
    L.pcsf_lpr.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp]
    L.pcsf_models_set.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp]
    L.pcsf_omega_models_set.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp]
    L.pcsf_model_get.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp]
    L.pcsf_pt_build_pairs.argtypes = [vp, i64, vp, vp, vp]
    L.pcsf_lpr_pairs.argtypes = [vp, i64, vp, vp, vp, vp, vp]
    L.pcsf_column_terms.argtypes = [vp, ctypes.c_int, vp, vp]
    L.pcsf_maximize_lpr.argtypes = [vp, ctypes.c_int, dbl, dbl, dbl, dbl, vp, vp, vp, vp, vp]
    L.pcsf_maximize_lpr_multi.argtypes = [vp, ctypes.c_int, vp, dbl, dbl, dbl, dbl, vp, vp, vp, vp, vp]
    L.pcsf_last_ms.argtypes = [vp, ctypes.c_int]
    L.pcsf_last_ms.restype = dbl
    L.pcsf_posteriors.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp]
    L.pcsf_counter.argtypes = [vp, ctypes.c_int]
    L.pcsf_counter.restype = i64
    L.pcsf_omega_cache_reset.argtypes = [vp, i64]
    L.pcsf_omega_models_set_cached.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp]
    L.pcsf_total_ms.argtypes = [vp, ctypes.c_int]
    L.pcsf_total_ms.restype = dbl
    L.pcsf_launch_count.argtypes = [vp]
    L.pcsf_launch_count.restype = i64
    L.pcsf_tree_n_leaves.argtypes = [vp]
    L.pcsf_table_level.argtypes = [vp, ctypes.c_int, ctypes.c_int]
    L.pcsf_last_launch_info.argtypes = [vp, ctypes.c_int]
    L.pcsf_last_launch_info.restype = i64
    _lib = L
    return L


def ptr(a):
    """Host pointer of a C-contiguous numpy array (or a raw integer address, or None)."""
    if a is None:
        return None
    if isinstance(a, int):
        return ctypes.c_void_p(a)
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"], "need a C-contiguous numpy array"
    return ctypes.c_void_p(a.ctypes.data)
