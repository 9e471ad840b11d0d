"""ctypes binding of libsassy_b200.so (the C ABI in include/sassy.h + include/sassy_gpu.h).

There is no CPU fallback: if the library is missing, or no B200 is visible when a
Searcher is constructed, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SASSY_B200_LIB selects another build of the same library (kernel tuning experiments).
LIB_PATH = os.environ.get("SASSY_B200_LIB") or os.path.join(HERE, "lib", "libsassy_b200.so")

c_size_t = ctypes.c_size_t
c_void_p = ctypes.c_void_p


class CMatch(ctypes.Structure):
    """sassy_Match, include/sassy.h (reference c/sassy.h:11-21)."""

    _fields_ = [
        ("text_start", ctypes.c_size_t),
        ("text_end", ctypes.c_size_t),
        ("pattern_start", ctypes.c_size_t),
        ("pattern_end", ctypes.c_size_t),
        ("cost", ctypes.c_int32),
        ("strand", ctypes.c_uint8),
    ]


class GpuMatch(ctypes.Structure):
    """sassy_gpu_Match, include/sassy_gpu.h."""

    _fields_ = [
        ("pattern_idx", ctypes.c_uint64),
        ("text_idx", ctypes.c_uint64),
        ("text_start", ctypes.c_uint64),
        ("text_end", ctypes.c_uint64),
        ("pattern_start", ctypes.c_uint64),
        ("pattern_end", ctypes.c_uint64),
        ("cost", ctypes.c_int32),
        ("strand", ctypes.c_uint8),
        ("reserved", ctypes.c_uint8 * 3),
        ("ops_len", ctypes.c_uint32),
        ("reserved2", ctypes.c_uint32),
        ("ops_off", ctypes.c_uint64),
    ]


class GpuStats(ctypes.Structure):
    """sassy_gpu_Stats, include/sassy_gpu.h."""

    _fields_ = [
        ("scan_ms", ctypes.c_float),
        ("total_ms", ctypes.c_float),
        ("scan_launches", ctypes.c_uint32),
        ("aux_launches", ctypes.c_uint32),
        ("candidates", ctypes.c_uint64),
        ("matches", ctypes.c_uint64),
        ("row_bytes", ctypes.c_uint32),
        ("rows", ctypes.c_uint32),
        ("words", ctypes.c_uint32),
        ("blocks_per_sm", ctypes.c_uint32),
        ("retries", ctypes.c_uint32),
        ("filter_words", ctypes.c_uint32),
        ("filter_ms", ctypes.c_float),
        ("verify_ms", ctypes.c_float),
        ("hits", ctypes.c_uint64),
        ("filter_len", ctypes.c_uint32),
        ("filter_fallback", ctypes.c_uint32),
        ("transfer_ms", ctypes.c_float),
        ("transfer_packed", ctypes.c_uint32),
        ("transfer_bytes", ctypes.c_uint64),
        ("filter_kind", ctypes.c_uint32),
        ("swar_lanes", ctypes.c_uint32),
        ("confirmed", ctypes.c_uint64),
        ("dense_tiles", ctypes.c_uint32),
        ("reserved3", ctypes.c_uint32),
    ]


# name -> (restype, argtypes); every symbol include/*.h declares.
SIGNATURES = {
    # include/sassy.h
    "sassy_searcher": (c_void_p, [ctypes.c_char_p, ctypes.c_bool, ctypes.c_float]),
    "sassy_searcher_free": (None, [c_void_p]),
    "search": (c_size_t, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_size_t,
                          ctypes.POINTER(ctypes.POINTER(CMatch))]),
    "sassy_matches_free": (None, [ctypes.POINTER(CMatch), c_size_t]),
    # include/sassy_gpu.h
    "sassy_gpu_device_count": (ctypes.c_int, []),
    "sassy_gpu_last_error": (ctypes.c_char_p, []),
    "sassy_gpu_device_info": (ctypes.c_int, [ctypes.c_int, ctypes.c_char_p, c_size_t, ctypes.POINTER(ctypes.c_int),
                                             ctypes.POINTER(ctypes.c_int), ctypes.POINTER(c_size_t),
                                             ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "sassy_gpu_searcher": (c_void_p, [ctypes.c_char_p, ctypes.c_bool, ctypes.c_float, ctypes.c_int]),
    "sassy_gpu_set_variant": (ctypes.c_int, [c_void_p, ctypes.c_int]),
    "sassy_gpu_set_filter": (ctypes.c_int, [c_void_p, ctypes.c_int]),
    "sassy_gpu_set_transport": (ctypes.c_int, [c_void_p, ctypes.c_int]),
    "sassy_gpu_stats": (ctypes.c_int, [c_void_p, ctypes.POINTER(GpuStats)]),
    "sassy_gpu_host_alloc": (c_void_p, [c_size_t]),
    "sassy_gpu_host_free": (None, [c_void_p]),
    "sassy_gpu_text_upload": (c_void_p, [c_void_p, c_void_p, c_size_t]),
    "sassy_gpu_text_from_device": (c_void_p, [c_void_p, c_void_p, c_size_t]),
    "sassy_gpu_text_len": (c_size_t, [c_void_p]),
    "sassy_gpu_text_free": (None, [c_void_p, c_void_p]),
    "sassy_gpu_search": (c_void_p, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, ctypes.c_int]),
    "sassy_gpu_search_text": (c_void_p, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, ctypes.c_int]),
    "sassy_gpu_set_trace": (ctypes.c_int, [c_void_p, ctypes.c_int]),
    "sassy_gpu_set_only_best_match": (ctypes.c_int, [c_void_p, ctypes.c_int]),
    "sassy_gpu_set_max_n_frac": (ctypes.c_int, [c_void_p, ctypes.c_float]),
    "sassy_gpu_set_max_overhang": (ctypes.c_int, [c_void_p, ctypes.c_int]),
    "sassy_gpu_search_pam": (c_void_p, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, ctypes.c_int,
                                        c_void_p, c_size_t]),
    "sassy_gpu_search_pam_text": (c_void_p, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, ctypes.c_int,
                                             c_void_p, c_size_t]),
    "sassy_gpu_search_patterns": (c_void_p, [c_void_p, c_void_p, c_size_t, c_size_t, c_void_p, c_size_t, c_size_t]),
    "sassy_gpu_search_texts": (c_void_p, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_size_t, c_size_t]),
    "sassy_gpu_search_many": (c_void_p, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_size_t,
                                         c_size_t]),
    "sassy_gpu_encode_patterns": (c_void_p, [c_void_p, c_void_p, c_size_t, c_size_t]),
    "sassy_gpu_patterns_free": (None, [c_void_p]),
    "sassy_gpu_search_encoded": (c_void_p, [c_void_p, c_void_p, c_void_p, c_size_t, ctypes.c_int]),
    "sassy_gpu_search_encoded_host": (c_void_p, [c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, ctypes.c_int]),
    "sassy_gpu_gather_create": (c_void_p, [c_void_p, ctypes.c_int, ctypes.c_int, c_size_t, c_size_t]),
    "sassy_gpu_gather_handle": (ctypes.c_int, [c_void_p, c_void_p]),
    "sassy_gpu_gather_connect": (ctypes.c_int, [c_void_p, c_void_p]),
    "sassy_gpu_gather_free": (None, [c_void_p]),
    "sassy_gpu_search_text_gathered": (c_void_p, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_size_t,
                                                  ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
    "sassy_gpu_search_encoded_gathered": (c_void_p, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, ctypes.c_int,
                                                     ctypes.POINTER(ctypes.c_int)]),
    "sassy_gpu_search_text_sharded": (c_void_p, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, ctypes.c_int,
                                                c_void_p, c_size_t, ctypes.c_uint64, ctypes.POINTER(ctypes.c_int)]),
    "sassy_gpu_gather_set_pipelined": (ctypes.c_int, [c_void_p, ctypes.c_int]),
    "sassy_gpu_gather_has_result": (ctypes.c_int, [c_void_p]),
    "sassy_gpu_text_sharded_flush": (c_void_p, [c_void_p, c_void_p, c_size_t, ctypes.c_int, c_void_p, c_size_t,
                                               ctypes.c_uint64, ctypes.POINTER(ctypes.c_int)]),
    "sassy_gpu_merge_slabs": (c_void_p, [c_void_p, c_size_t, c_void_p, c_void_p, c_size_t, ctypes.c_uint64,
                                        ctypes.c_int]),
    "sassy_gpu_result_len": (c_size_t, [c_void_p]),
    "sassy_gpu_result_matches": (ctypes.POINTER(GpuMatch), [c_void_p]),
    "sassy_gpu_result_ops": (c_void_p, [c_void_p]),
    "sassy_gpu_cigar": (c_size_t, [c_void_p, c_size_t, ctypes.c_char_p, c_size_t]),
    "sassy_gpu_result_free": (None, [c_void_p]),
}

_LIB = None


def load() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m sassy_b200.build` "
                "(needs nvcc; there is no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_LOCAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def last_error() -> str:
    msg = load().sassy_gpu_last_error()
    return msg.decode(errors="replace") if msg else ""
