"""ctypes loader for libgranges_b200.so (the C ABI declared in include/granges_b200.h).

There is no fallback: if the shared library is missing or cannot be loaded the import fails
loudly -- the product path never routes through a CPU implementation.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GRB_LIB_PATH") or os.path.join(_HERE, "libgranges_b200.so")  # the override is for A/B builds of the same ABI

u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)
u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int)
vp = C.c_void_p


class Shard(C.Structure):
    _fields_ = [("seq", C.c_int32), ("rank", C.c_int32), ("left_begin", C.c_uint64), ("left_end", C.c_uint64),
                ("right_lo", C.c_uint32), ("right_hi", C.c_uint32)]


# name -> (restype, argtypes); every symbol include/granges_b200.h declares
SIGNATURES = {
    "grb_abi_version": (C.c_int, []),
    "grb_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "grb_ctx_set_stream": (C.c_int, [vp, vp]),
    "grb_ctx_use_own_stream": (C.c_int, [vp]),
    "grb_ctx_sync": (C.c_int, [vp]),
    "grb_ctx_set_left_order": (C.c_int, [vp, C.c_int]),
    "grb_ctx_destroy": (None, [vp]),
    "grb_last_error": (C.c_char_p, [vp]),
    "grb_ctx_launch_count": (C.c_uint64, [vp]),
    "grb_ctx_sm_count": (C.c_int, [vp]),
    "grb_index_build": (C.c_int, [vp, vp, vp, vp, C.c_size_t, C.POINTER(vp)]),
    "grb_index_build_batch": (C.c_int, [vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(vp), C.POINTER(vp),
                                         C.POINTER(vp)]),
    "grb_index_destroy": (None, [vp]),
    "grb_index_len": (C.c_size_t, [vp]),
    "grb_index_set_scores": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "grb_index_copy_sorted": (C.c_int, [vp, vp, vp, vp]),
    "grb_index_exact_sums": (C.c_int, [vp]),
    "grb_count_overlaps": (C.c_int, [vp, vp, vp, vp, C.c_size_t, vp]),
    "grb_filter_overlaps": (C.c_int, [vp, vp, vp, vp, C.c_size_t, C.c_int, vp, u64p]),
    "grb_left_overlaps": (C.c_int, [vp, vp, vp, vp, C.c_size_t, vp, C.POINTER(vp)]),
    "grb_pairs_len": (C.c_size_t, [vp]),
    "grb_pairs_copy": (C.c_int, [vp, vp, vp, vp]),
    "grb_pairs_overlap_widths": (C.c_int, [vp, vp]),
    "grb_pairs_destroy": (None, [vp]),
    "grb_map_joins_f64": (C.c_int, [vp, vp, vp, vp, C.c_size_t, i32p, C.c_int, vp, vp]),
    "grb_count_overlaps_batch": (C.c_int, [vp, C.POINTER(vp), u64p, C.c_int, vp, vp, vp]),
    "grb_filter_overlaps_batch": (C.c_int, [vp, C.POINTER(vp), u64p, C.c_int, vp, vp, C.c_int, vp, u64p]),
    "grb_filter_overlaps_select": (C.c_int, [vp, C.POINTER(vp), u64p, C.c_int, vp, vp, C.c_int, vp, u64p]),
    "grb_map_joins_f64_batch": (C.c_int, [vp, C.POINTER(vp), u64p, C.c_int, vp, vp, i32p, C.c_int, vp, vp]),
    "grb_map_joins_f64_batch_bits": (C.c_int, [vp, C.POINTER(vp), u64p, C.c_int, vp, vp, i32p, C.c_int, vp, vp]),
    "grb_map_joins_reduce_batch": (C.c_int, [vp, C.POINTER(vp), u64p, C.c_int, vp, vp, vp, C.c_int, vp, vp]),
    "grb_collapse_f64": (C.c_int, [vp, vp, vp, vp, C.c_size_t, C.POINTER(vp)]),
    "grb_text_bytes": (C.c_size_t, [vp]),
    "grb_text_copy": (C.c_int, [vp, vp, vp]),
    "grb_text_destroy": (None, [vp]),
    "grb_map_joins_whole_f64": (C.c_int, [vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(vp), C.POINTER(vp),
                                           u64p, vp, vp, i32p, C.c_int, vp, vp]),
    "grb_map_joins_whole_f64_mgpu": (C.c_int, [i32p, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(vp),
                                                C.POINTER(vp), u64p, vp, vp, i32p, C.c_int, vp, vp, C.c_char_p, C.c_size_t]),
    "grb_windows": (C.c_int, [vp, vp, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(vp)]),
    "grb_ranges_len": (C.c_size_t, [vp]),
    "grb_ranges_seq_offsets": (C.c_int, [vp, u64p]),
    "grb_ranges_copy": (C.c_int, [vp, vp, vp, vp]),
    "grb_ranges_dev_starts": (vp, [vp]),
    "grb_ranges_dev_ends": (vp, [vp]),
    "grb_ranges_destroy": (None, [vp]),
    "grb_merge_ranges": (C.c_int, [vp, vp, vp, u64p, C.c_int, C.c_int64, C.POINTER(vp)]),
    "grb_merged_len": (C.c_size_t, [vp]),
    "grb_merged_copy": (C.c_int, [vp, vp, vp, vp, u64p]),
    "grb_merged_dev_starts": (vp, [vp]),
    "grb_merged_dev_ends": (vp, [vp]),
    "grb_merged_map_f64": (C.c_int, [vp, vp, vp, vp, i32p, C.c_int, vp, vp]),
    "grb_merged_destroy": (None, [vp]),
    "grb_shard_plan": (C.c_int, [u64p, u64p, C.c_int, vp, vp, C.c_int, C.POINTER(Shard), C.c_int]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C granges_b200/csrc`). granges_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
