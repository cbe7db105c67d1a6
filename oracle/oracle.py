"""ctypes binding of oracle/granges_oracle.c (the coitrees-restated CPU oracle).

TEST INFRASTRUCTURE ONLY -- see oracle/granges_oracle.c for what it restates
(reference file:line citations live there).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

OP_SUM, OP_SUM_NOT_EMPTY, OP_MIN, OP_MAX, OP_MEAN, OP_MEDIAN, OP_COUNT, OP_OVERLAP_BP, OP_WEIGHTED_MEAN, OP_WEIGHTED_SUM = range(10)
OPS = {
    "sum": OP_SUM,
    "sum-not-empty": OP_SUM_NOT_EMPTY,
    "min": OP_MIN,
    "max": OP_MAX,
    "mean": OP_MEAN,
    "median": OP_MEDIAN,
    "count": OP_COUNT,
    "overlap-bp": OP_OVERLAP_BP,
    "weighted-mean": OP_WEIGHTED_MEAN,
    "weighted-sum": OP_WEIGHTED_SUM,
}


def build(force: bool = False) -> str:
    """Compile liboracle.so with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "granges_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u32p, u64p, f64p, u8p, i32p = (C.POINTER(t) for t in (C.c_uint32, C.c_uint64, C.c_double, C.c_uint8, C.c_int))
        L.orc_tree_build.restype = C.c_void_p
        L.orc_tree_build.argtypes = [u32p, u32p, u64p, C.c_size_t]
        L.orc_tree_free.argtypes = [C.c_void_p]
        L.orc_tree_len.restype = C.c_size_t
        L.orc_tree_len.argtypes = [C.c_void_p]
        L.orc_count_overlaps.restype = C.c_size_t
        L.orc_count_overlaps.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.orc_query.restype = C.c_size_t
        L.orc_query.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, u64p, u32p, u32p, C.c_size_t]
        L.orc_left_overlaps.restype = C.c_int
        L.orc_left_overlaps.argtypes = [C.c_void_p, u32p, u32p, C.c_size_t, u64p, C.POINTER(u64p), C.POINTER(u32p), C.POINTER(u32p)]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_overlap_width.restype = C.c_uint32
        L.orc_overlap_width.argtypes = [C.c_uint32] * 4
        L.orc_median.restype = C.c_int
        L.orc_median.argtypes = [f64p, C.c_size_t, f64p]
        L.orc_float_op.restype = C.c_int
        L.orc_float_op.argtypes = [C.c_int, f64p, C.c_size_t, f64p]
        L.orc_map_joins_f64.restype = C.c_int
        L.orc_map_joins_f64.argtypes = [C.c_void_p, f64p, u8p, u32p, u32p, C.c_size_t, i32p, C.c_int, f64p, u8p, C.c_int]
        L.orc_filter_overlaps.restype = C.c_int
        L.orc_filter_overlaps.argtypes = [C.c_void_p, u32p, u32p, C.c_size_t, C.c_int, u8p]
        L.orc_windows.restype = C.c_size_t
        L.orc_windows.argtypes = [u32p, C.c_int, C.c_uint32, C.c_uint32, C.c_int, u32p, u32p, u32p, C.c_size_t]
        L.orc_merge.restype = C.c_size_t
        L.orc_merge.argtypes = [u32p, u32p, u32p, C.c_size_t, C.c_int64, f64p, u8p, C.c_int, u32p, u32p, u32p, u64p, f64p, u8p]
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


class Tree:
    """One seqname's COITrees (ranges/coitrees.rs:25-84)."""

    def __init__(self, starts, ends, idx=None):
        self.starts = _u32(starts)
        self.ends = _u32(ends)
        self.idx = None if idx is None else np.ascontiguousarray(idx, dtype=np.uint64)
        n = len(self.starts)
        self.h = lib().orc_tree_build(_p(self.starts, C.c_uint32), _p(self.ends, C.c_uint32), _p(self.idx, C.c_uint64), n)
        if not self.h:
            raise ValueError("invalid range (end <= start or coordinate >= 2^31)")

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_tree_free(self.h)
            self.h = None

    def __len__(self):
        return lib().orc_tree_len(self.h)

    def count_overlaps(self, start, end):
        return lib().orc_count_overlaps(self.h, start, end)

    def query(self, start, end):
        n = lib().orc_query(self.h, start, end, None, None, None, 0)
        idx = np.empty(n, np.uint64)
        rs = np.empty(n, np.uint32)
        re = np.empty(n, np.uint32)
        lib().orc_query(self.h, start, end, _p(idx, C.c_uint64), _p(rs, C.c_uint32), _p(re, C.c_uint32), n)
        return idx, rs, re


def _h(tree):
    return None if tree is None else tree.h


def left_overlaps(tree, ls, le):
    """CSR (offsets[nl+1], right_idx[P], r_start[P], r_end[P]); groups in sort_ranges order."""
    ls, le = _u32(ls), _u32(le)
    nl = len(ls)
    offsets = np.zeros(nl + 1, np.uint64)
    pi, ps, pe = C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)()
    lib().orc_left_overlaps(_h(tree), _p(ls, C.c_uint32), _p(le, C.c_uint32), nl, _p(offsets, C.c_uint64), C.byref(pi), C.byref(ps), C.byref(pe))
    P = int(offsets[-1])
    idx = np.ctypeslib.as_array(pi, (P,)).copy() if P else np.empty(0, np.uint64)
    rs = np.ctypeslib.as_array(ps, (P,)).copy() if P else np.empty(0, np.uint32)
    re = np.ctypeslib.as_array(pe, (P,)).copy() if P else np.empty(0, np.uint32)
    for p in (pi, ps, pe):
        lib().orc_free(C.cast(p, C.c_void_p))
    return offsets, idx, rs, re


def map_joins(tree, scores, valid, ls, le, ops, nthreads=1):
    """(out[nl,k] f64, out_valid[nl,k] u8) -- left_overlaps + map_joins(FloatOperation...)."""
    ls, le = _u32(ls), _u32(le)
    nl = len(ls)
    ops_a = np.ascontiguousarray([OPS[o] if isinstance(o, str) else int(o) for o in ops], dtype=np.int32)
    k = len(ops_a)
    sc = None if scores is None else np.ascontiguousarray(scores, dtype=np.float64)
    va = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8)
    out = np.zeros((nl, k), np.float64)
    ov = np.zeros((nl, k), np.uint8)
    rc = lib().orc_map_joins_f64(_h(tree), _p(sc, C.c_double), _p(va, C.c_uint8), _p(ls, C.c_uint32), _p(le, C.c_uint32), nl, _p(ops_a, C.c_int), k, _p(out, C.c_double), _p(ov, C.c_uint8), nthreads)
    if rc != 0:
        raise ValueError("bad op")
    return out, ov


def filter_overlaps(tree, ls, le, anti=False):
    ls, le = _u32(ls), _u32(le)
    keep = np.zeros(len(ls), np.uint8)
    lib().orc_filter_overlaps(_h(tree), _p(ls, C.c_uint32), _p(le, C.c_uint32), len(ls), int(anti), _p(keep, C.c_uint8))
    return keep


def windows(seqlens, width, step=None, chop=False):
    """(seq_id, start, end) arrays -- GRangesEmpty::from_windows (granges.rs:634-666)."""
    sl = _u32(seqlens)
    st = 0 if step is None else int(step)
    n = lib().orc_windows(_p(sl, C.c_uint32), len(sl), width, st, int(chop), None, None, None, 0)
    seq = np.empty(n, np.uint32)
    s = np.empty(n, np.uint32)
    e = np.empty(n, np.uint32)
    lib().orc_windows(_p(sl, C.c_uint32), len(sl), width, st, int(chop), _p(seq, C.c_uint32), _p(s, C.c_uint32), _p(e, C.c_uint32), n)
    return seq, s, e


def overlap_width(a, b):
    return lib().orc_overlap_width(a[0], a[1], b[0], b[1])


def median(values):
    v = np.array(values, dtype=np.float64)
    out = C.c_double(0)
    ok = lib().orc_median(_p(v, C.c_double), len(v), C.byref(out))
    return out.value if ok else None


def float_op(op, values):
    v = np.array(values, dtype=np.float64)
    out = C.c_double(0)
    ok = lib().orc_float_op(OPS[op] if isinstance(op, str) else op, _p(v, C.c_double), len(v), C.byref(out))
    return out.value if ok == 1 else None


def merge(seq, starts, ends, distance=0, scores=None, valid=None, op=None):
    """Streaming merge of a file's records (merging_iterators.rs): (seq, start, end, first[, value, ok])."""
    seq, starts, ends = _u32(seq), _u32(starts), _u32(ends)
    n = len(starts)
    sc = None if scores is None else np.ascontiguousarray(scores, dtype=np.float64)
    va = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8)
    o_seq, o_s, o_e = np.empty(n, np.uint32), np.empty(n, np.uint32), np.empty(n, np.uint32)
    o_first, o_val, o_ok = np.empty(n, np.uint64), np.zeros(n, np.float64), np.zeros(n, np.uint8)
    opc = 0 if op is None else (OPS[op] if isinstance(op, str) else int(op))
    m = lib().orc_merge(_p(seq, C.c_uint32), _p(starts, C.c_uint32), _p(ends, C.c_uint32), n, int(distance), _p(sc, C.c_double),
                        _p(va, C.c_uint8), opc, _p(o_seq, C.c_uint32), _p(o_s, C.c_uint32), _p(o_e, C.c_uint32), _p(o_first, C.c_uint64),
                        _p(o_val, C.c_double), _p(o_ok, C.c_uint8))
    res = (o_seq[:m], o_s[:m], o_e[:m], o_first[:m])
    if scores is not None:
        res = res + (o_val[:m], o_ok[:m])
    return res


def max_threads():
    return lib().orc_max_threads()
