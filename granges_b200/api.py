"""Python face of the C ABI: thin, typed, numpy / torch friendly.

Array arguments may be numpy arrays (host) or torch CUDA tensors (device; passed by data_ptr,
results stay on the device).  Names follow the reference (vsbuffalo/granges v0.4.0):
Index ~ COITrees<M> (src/ranges/coitrees.rs), Context.left_overlaps ~ LeftOverlaps::left_overlaps
(src/traits.rs:44-48), Context.map_joins ~ left_overlaps + map_joins with FloatOperation reducers
(src/granges.rs:1176-1190, src/data/operations.rs:53-109), Context.filter_overlaps
(src/granges.rs:1416-1609), Context.windows ~ GRangesEmpty::from_windows (src/granges.rs:634-666).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi

OPS = {
    "sum": 0,
    "sum-not-empty": 1,
    "min": 2,
    "max": 3,
    "mean": 4,
    "median": 5,
    "count": 6,
    "overlap-bp": 7,
    "weighted-mean": 8,
    "weighted-sum": 9,
}

# grb_reducer parts (include/granges_b200.h)
RED_VALUE = {"score": 0, "one": 1}
RED_WEIGHT = {"one": 0, "overlap-width": 1}
RED_FOLD = {"sum": 0, "min": 1, "max": 2}
RED_DIVIDE = {"none": 0, "by-weights": 1}
RED_EMPTY = {"zero": 0, "novalue": 1}

STATUS = {
    0: "GRB_OK",
    -1: "GRB_ERR_INVALID_ARG",
    -2: "GRB_ERR_INVALID_RANGE",
    -3: "GRB_ERR_NO_DATA",
    -4: "GRB_ERR_CUDA",
    -5: "GRB_ERR_OOM",
    -6: "GRB_ERR_NAN_SCORE",
    -7: "GRB_ERR_INDEX_RANGE",
    -8: "GRB_ERR_UNSUPPORTED",
}


class GRangesError(RuntimeError):
    """Mirrors GRangesError (src/error.rs:72-176) / the reference's panics; .code is the grb_status."""

    def __init__(self, code, msg):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code


def _is_torch(a):
    return type(a).__module__.startswith("torch")


def _arg(a, dtype):
    """-> (pointer value or None, keep-alive object, is_device)."""
    if a is None:
        return None, None, False
    if _is_torch(a):
        import torch

        want = {np.uint32: (torch.uint32, torch.int32), np.uint64: (torch.uint64, torch.int64), np.float64: (torch.float64,),
                np.uint8: (torch.uint8, torch.bool)}[dtype]
        if a.dtype not in want:
            raise TypeError(f"tensor dtype {a.dtype} does not match {dtype.__name__}")
        if not a.is_contiguous():
            a = a.contiguous()
        return a.data_ptr(), a, a.is_cuda
    arr = np.ascontiguousarray(a, dtype=dtype)
    return arr.ctypes.data, arr, False


def _out_arg(a, dtype, shape=None):
    """Pointer of a caller-provided OUTPUT buffer.  Never copies: the buffer must already be contiguous, writeable,
    of the exact dtype (and shape, when given) -- a silently made temporary would receive the results instead of the
    caller's array.  -> (pointer, the buffer itself as keep-alive)."""
    if _is_torch(a):
        import torch

        want = {np.uint32: (torch.uint32, torch.int32), np.uint64: (torch.uint64, torch.int64), np.float64: (torch.float64,),
                np.uint8: (torch.uint8, torch.bool)}[dtype]
        if a.dtype not in want:
            raise TypeError(f"output tensor dtype {a.dtype} does not match {dtype.__name__}")
        if not a.is_contiguous():
            raise ValueError("output tensor must be contiguous")
        if shape is not None and tuple(a.shape) != tuple(shape):
            raise ValueError(f"output tensor has shape {tuple(a.shape)}, expected {tuple(shape)}")
        return a.data_ptr(), a
    if not isinstance(a, np.ndarray):
        raise TypeError("output buffer must be a numpy array or a torch tensor")
    if a.dtype != np.dtype(dtype):
        raise TypeError(f"output array dtype {a.dtype} does not match {np.dtype(dtype)}")
    if not a.flags.c_contiguous or not a.flags.writeable:
        raise ValueError("output array must be C-contiguous and writeable")
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"output array has shape {a.shape}, expected {tuple(shape)}")
    return a.ctypes.data, a


class Index:
    """One seqname's right-hand index (replaces COITrees<M>, src/ranges/coitrees.rs:25-84)."""

    def __init__(self, ctx, handle):
        self.ctx = ctx  # (keeps the context alive: a handle must never outlive its context)
        self.h = handle
        ctx._children.add(self)

    def __len__(self):
        return _capi.lib().grb_index_len(self.h)

    def set_scores(self, scores, valid=None):
        ps, ks, _ = _arg(scores, np.float64)
        pv, kv, _ = _arg(valid, np.uint8)
        n = len(ks) if ks is not None else 0
        self.ctx._check(_capi.lib().grb_index_set_scores(self.h, ps, pv, n))
        return self

    @property
    def exact_sums(self):
        return bool(_capi.lib().grb_index_exact_sums(self.h))

    def sorted_ranges(self):
        """COITrees::iter_ranges order (src/ranges/coitrees.rs:154-170): (starts, ends, data_idx)."""
        n = len(self)
        s, e, d = np.empty(n, np.uint32), np.empty(n, np.uint32), np.empty(n, np.uint64)
        self.ctx._check(_capi.lib().grb_index_copy_sorted(self.h, s.ctypes.data, e.ctypes.data, d.ctypes.data))
        return s, e, d

    def close(self):
        if self.h:
            _capi.lib().grb_index_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Windows:
    """Device-resident result of Context.windows_dev."""

    def __init__(self, ctx, handle, nseq):
        self.ctx, self.h, self.nseq = ctx, handle, nseq
        ctx._children.add(self)

    def __len__(self):
        return _capi.lib().grb_ranges_len(self.h)

    def seq_offsets(self):
        off = np.zeros(self.nseq + 1, np.uint64)
        self.ctx._check(_capi.lib().grb_ranges_seq_offsets(self.h, off.ctypes.data_as(_capi.u64p)))
        return off

    def dev_ptrs(self):
        L = _capi.lib()
        return L.grb_ranges_dev_starts(self.h), L.grb_ranges_dev_ends(self.h)

    def copy(self):
        n = len(self)
        q, s, e = np.empty(n, np.uint32), np.empty(n, np.uint32), np.empty(n, np.uint32)
        self.ctx._check(_capi.lib().grb_ranges_copy(self.h, q.ctypes.data, s.ctypes.data, e.ctypes.data))
        return q, s, e

    def close(self):
        if self.h:
            _capi.lib().grb_ranges_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One CUDA device + stream (grb_ctx)."""

    def __init__(self, device=0, stream=None):
        L = _capi.lib()
        h = C.c_void_p()
        rc = L.grb_ctx_create(int(device), C.byref(h))
        if rc != 0:
            raise GRangesError(rc, f"cannot create a context on CUDA device {device} (no CPU fallback exists)")
        self.h = h
        self.device = int(device)
        import weakref

        self._children = weakref.WeakSet()  # indexes / windows created here: closed before the context goes
        if stream is None:
            # Device tensors handed to this context are produced on torch's current stream: run there too, so that the
            # library's kernels are ordered after the producer and before the consumer of every tensor (the context's own
            # non-blocking stream would race with both).  Callers who manage streams themselves pass one explicitly.
            import sys

            torch = sys.modules.get("torch")
            if torch is not None and torch.cuda.is_available():
                stream = torch.cuda.current_stream(self.device)
        if stream is not None:
            self.set_stream(stream)

    # -- plumbing
    def _check(self, rc):
        if rc != 0:
            raise GRangesError(rc, _capi.lib().grb_last_error(self.h).decode())

    def set_stream(self, stream):
        """Run on `stream` (a torch.cuda.Stream or a raw cudaStream_t; 0 is the default stream); None = own stream."""
        if stream is None:
            self._check(_capi.lib().grb_ctx_use_own_stream(self.h))
            return
        ptr = int(getattr(stream, "cuda_stream", stream))
        self._check(_capi.lib().grb_ctx_set_stream(self.h, C.c_void_p(ptr)))

    def set_left_order(self, order):
        """"auto" (default) | "sorted" (never bucket) | "unsorted" (always bucket): grb_ctx_set_left_order."""
        code = {"auto": 0, "sorted": 1, "unsorted": 2}[order] if isinstance(order, str) else int(order)
        self._check(_capi.lib().grb_ctx_set_left_order(self.h, code))

    def sync(self):
        self._check(_capi.lib().grb_ctx_sync(self.h))

    @property
    def launch_count(self):
        return int(_capi.lib().grb_ctx_launch_count(self.h))

    @property
    def sm_count(self):
        return int(_capi.lib().grb_ctx_sm_count(self.h))

    def close(self):
        if self.h:
            for ch in list(self._children):
                ch.close()
            _capi.lib().grb_ctx_destroy(self.h)
            self.h = None

    # -- index
    def index_build(self, starts, ends, data_idx=None):
        ps, ks, _ = _arg(starts, np.uint32)
        pe, ke, _ = _arg(ends, np.uint32)
        pi, ki, _ = _arg(data_idx, np.uint64)
        n = len(ks)
        if len(ke) != n or (ki is not None and len(ki) != n):
            raise ValueError("starts / ends / data_idx lengths differ")
        h = C.c_void_p()
        self._check(_capi.lib().grb_index_build(self.h, ps, pe, pi, n, C.byref(h)))
        return Index(self, h)

    def index_build_batch(self, rights):
        """Indexes of a whole right-hand GRanges: `rights` = [(starts, ends[, scores[, valid]]) per seqname].
        Host-to-device copies of one seqname overlap the indexing of the previous one."""
        nseq = len(rights)
        keep, ps, pe, psc, pva, ns = [], [], [], [], [], []
        any_scores = any(len(r) > 2 and r[2] is not None for r in rights)
        any_valid = any(len(r) > 3 and r[3] is not None for r in rights)
        for r in rights:
            a, ka, _ = _arg(r[0], np.uint32)
            b, kb, _ = _arg(r[1], np.uint32)
            c, kc, _ = _arg(r[2] if len(r) > 2 else None, np.float64)
            d, kd, _ = _arg(r[3] if len(r) > 3 else None, np.uint8)
            keep += [ka, kb, kc, kd]
            ps.append(a)
            pe.append(b)
            psc.append(c)
            pva.append(d)
            ns.append(len(ka))
        VP = C.c_void_p * nseq
        out = VP()
        self._check(_capi.lib().grb_index_build_batch(self.h, nseq, VP(*ps), VP(*pe), (C.c_size_t * nseq)(*ns),
                                                      VP(*psc) if any_scores else None, VP(*pva) if any_valid else None, out))
        return [Index(self, C.c_void_p(h)) for h in out]

    def map_whole(self, rights, left_offsets, ls, le, ops, out=None, out_valid=None):
        """`granges map` in one call for host arrays (grb_map_joins_whole_f64): `rights` as in index_build_batch,
        lefts concatenated in seqname order.  Uploads, index builds, map kernels and downloads of different seqnames
        overlap; the results equal index_build_batch + map_joins_batch."""
        nseq = len(rights)
        keep, ps, pe, psc, pva, ns = [], [], [], [], [], []
        any_scores = any(len(r) > 2 and r[2] is not None for r in rights)
        any_valid = any(len(r) > 3 and r[3] is not None for r in rights)
        for r in rights:
            a, ka, _ = _arg(r[0], np.uint32)
            b, kb, _ = _arg(r[1], np.uint32)
            c, kc, _ = _arg(r[2] if len(r) > 2 else None, np.float64)
            d, kd, _ = _arg(r[3] if len(r) > 3 else None, np.uint8)
            keep += [ka, kb, kc, kd]
            ps.append(a)
            pe.append(b)
            psc.append(c)
            pva.append(d)
            ns.append(len(ka))
        pl, kl, _ = _arg(ls, np.uint32)
        ple, kle, _ = _arg(le, np.uint32)
        off = np.ascontiguousarray(left_offsets, dtype=np.uint64)
        ops_a, ops_p = self._ops(ops)
        k = len(ops_a)
        nl = len(kl)
        if out is None:
            out = np.empty((nl, k), np.float64)
        if out_valid is None:
            out_valid = np.empty((nl, k), np.uint8)
        po, pv = _out_arg(out, np.float64, (nl, k))[0], _out_arg(out_valid, np.uint8, (nl, k))[0]
        VP = C.c_void_p * nseq
        self._check(_capi.lib().grb_map_joins_whole_f64(self.h, nseq, VP(*ps), VP(*pe), (C.c_size_t * nseq)(*ns),
                                                        VP(*psc) if any_scores else None, VP(*pva) if any_valid else None,
                                                        off.ctypes.data_as(_capi.u64p), pl, ple, ops_p, k, po, pv))
        return out, out_valid

    # -- helpers
    @staticmethod
    def _ops(ops):
        a = np.ascontiguousarray([OPS[o] if isinstance(o, str) else int(o) for o in ops], dtype=np.int32)
        return a, a.ctypes.data_as(_capi.i32p)

    def _out(self, like_device, shape, dtype, like=None):
        if like_device:
            import torch

            tdt = {np.uint32: torch.int32, np.uint8: torch.uint8, np.float64: torch.float64, np.uint64: torch.int64}[dtype]
            t = torch.empty(shape, dtype=tdt, device=like.device)
            return t.data_ptr(), t
        a = np.empty(shape, dtype)
        return a.ctypes.data, a

    # -- queries (single seqname)
    def count_overlaps(self, index, ls, le):
        pl, kl, dev = _arg(ls, np.uint32)
        pe, ke, _ = _arg(le, np.uint32)
        nl = len(kl)
        po, out = self._out(dev, (nl,), np.uint32, kl)
        self._check(_capi.lib().grb_count_overlaps(self.h, index.h if index is not None else None, pl, pe, nl, po))
        return out

    def filter_overlaps(self, index, ls, le, anti=False, return_count=False):
        pl, kl, dev = _arg(ls, np.uint32)
        pe, ke, _ = _arg(le, np.uint32)
        nl = len(kl)
        po, out = self._out(dev, (nl,), np.uint8, kl)
        nk = C.c_uint64(0)
        self._check(_capi.lib().grb_filter_overlaps(self.h, index.h if index is not None else None, pl, pe, nl, int(bool(anti)), po,
                                                    C.byref(nk)))
        return (out, nk.value) if return_count else out

    def left_overlaps(self, index, ls, le, widths=False):
        """CSR join: (offsets[nl+1] u64, right_data_idx[P] u64, r_start[P], r_end[P][, overlap_widths[P]])."""
        L = _capi.lib()
        pl, kl, _ = _arg(ls, np.uint32)
        pe, ke, _ = _arg(le, np.uint32)
        nl = len(kl)
        offsets = np.zeros(nl + 1, np.uint64)
        ph = C.c_void_p()
        self._check(L.grb_left_overlaps(self.h, index.h if index is not None else None, pl, pe, nl, offsets.ctypes.data, C.byref(ph)))
        try:
            P = L.grb_pairs_len(ph)
            idx, rs, re = np.empty(P, np.uint64), np.empty(P, np.uint32), np.empty(P, np.uint32)
            self._check(L.grb_pairs_copy(ph, idx.ctypes.data, rs.ctypes.data, re.ctypes.data))
            res = (offsets, idx, rs, re)
            if widths:
                w = np.empty(P, np.uint32)
                self._check(L.grb_pairs_overlap_widths(ph, w.ctypes.data))
                res = res + (w,)
        finally:
            L.grb_pairs_destroy(ph)
        return res

    def map_joins(self, index, ls, le, ops):
        pl, kl, dev = _arg(ls, np.uint32)
        pe, ke, _ = _arg(le, np.uint32)
        nl = len(kl)
        ops_a, ops_p = self._ops(ops)
        k = len(ops_a)
        po, out = self._out(dev, (nl, k), np.float64, kl)
        pv, ov = self._out(dev, (nl, k), np.uint8, kl)
        self._check(_capi.lib().grb_map_joins_f64(self.h, index.h if index is not None else None, pl, pe, nl, ops_p, k, po, pv))
        return out, ov

    # -- queries (whole GRanges: seqnames concatenated in GenomeMap order)
    @staticmethod
    def _handles(indexes):
        arr = (C.c_void_p * len(indexes))(*[(ix.h if ix is not None else None) for ix in indexes])
        return arr

    def count_overlaps_batch(self, indexes, left_offsets, ls, le, out=None):
        pl, kl, dev = _arg(ls, np.uint32)
        pe, ke, _ = _arg(le, np.uint32)
        off = np.ascontiguousarray(left_offsets, dtype=np.uint64)
        if out is None:
            po, out = self._out(dev, (len(kl),), np.uint32, kl)
        else:
            po = _out_arg(out, np.uint32, (len(kl),))[0]
        self._check(_capi.lib().grb_count_overlaps_batch(self.h, self._handles(indexes), off.ctypes.data_as(_capi.u64p), len(indexes),
                                                         pl, pe, po))
        return out

    def filter_overlaps_batch(self, indexes, left_offsets, ls, le, anti=False, out=None):
        pl, kl, dev = _arg(ls, np.uint32)
        pe, ke, _ = _arg(le, np.uint32)
        off = np.ascontiguousarray(left_offsets, dtype=np.uint64)
        if out is None:
            po, out = self._out(dev, (len(kl),), np.uint8, kl)
        else:
            po = _out_arg(out, np.uint8, (len(kl),))[0]
        nk = C.c_uint64(0)
        self._check(_capi.lib().grb_filter_overlaps_batch(self.h, self._handles(indexes), off.ctypes.data_as(_capi.u64p), len(indexes),
                                                          pl, pe, int(bool(anti)), po, C.byref(nk)))
        return out, nk.value

    def filter_overlaps_select(self, indexes, left_offsets, ls, le, anti=False):
        """Ascending indices of the lefts the semi- (anti-) join keeps: the rows of the filtered GRanges."""
        pl, kl, dev = _arg(ls, np.uint32)
        pe, ke, _ = _arg(le, np.uint32)
        off = np.ascontiguousarray(left_offsets, dtype=np.uint64)
        po, out = self._out(dev, (len(kl),), np.uint64, kl)
        nk = C.c_uint64(0)
        self._check(_capi.lib().grb_filter_overlaps_select(self.h, self._handles(indexes), off.ctypes.data_as(_capi.u64p), len(indexes),
                                                           pl, pe, int(bool(anti)), po, C.byref(nk)))
        return out[: nk.value]

    def map_joins_batch(self, indexes, left_offsets, ls, le, ops, out=None, out_valid=None, valid_bits=False):
        """valid_bits=True: the validity comes back as a u32 bit array (bit i * k + c), grb_map_joins_f64_batch_bits."""
        pl, kl, dev = _arg(ls, np.uint32)
        pe, ke, _ = _arg(le, np.uint32)
        off = np.ascontiguousarray(left_offsets, dtype=np.uint64)
        ops_a, ops_p = self._ops(ops)
        k = len(ops_a)
        nl = len(kl)
        vshape, vdt = (((nl * k + 31) // 32,), np.uint32) if valid_bits else ((nl, k), np.uint8)
        if out is None:
            po, out = self._out(dev, (nl, k), np.float64, kl)
        else:
            po = _out_arg(out, np.float64, (nl, k))[0]
        if out_valid is None:
            pv, out_valid = self._out(dev, vshape, vdt, kl)
        else:
            pv = _out_arg(out_valid, vdt, vshape)[0]
        fn = _capi.lib().grb_map_joins_f64_batch_bits if valid_bits else _capi.lib().grb_map_joins_f64_batch
        self._check(fn(self.h, self._handles(indexes), off.ctypes.data_as(_capi.u64p), len(indexes), pl, pe, ops_p, k, po, pv))
        return out, out_valid

    def map_joins_reduce(self, indexes, left_offsets, ls, le, reducers):
        """Composable reducers (grb_map_joins_reduce_batch): `reducers` = [(value, weight, fold, divide, empty), ...] with the
        names of RED_VALUE / RED_WEIGHT / RED_FOLD / RED_DIVIDE / RED_EMPTY, one per output column."""
        pl, kl, dev = _arg(ls, np.uint32)
        pe, ke, _ = _arg(le, np.uint32)
        off = np.ascontiguousarray(left_offsets, dtype=np.uint64)
        k = len(reducers)
        red = np.ascontiguousarray([[RED_VALUE[v], RED_WEIGHT[w], RED_FOLD[f], RED_DIVIDE[d], RED_EMPTY[e]] for v, w, f, d, e in reducers],
                                   dtype=np.int32)
        po, out = self._out(dev, (len(kl), k), np.float64, kl)
        pv, ov = self._out(dev, (len(kl), k), np.uint8, kl)
        self._check(_capi.lib().grb_map_joins_reduce_batch(self.h, self._handles(indexes), off.ctypes.data_as(_capi.u64p), len(indexes), pl, pe,
                                                           red.ctypes.data, k, po, pv))
        return out, ov

    def collapse(self, index, ls, le):
        """FloatOperation::Collapse in sort_ranges order (grb_collapse_f64): one string per left."""
        L = _capi.lib()
        pl, kl, _ = _arg(ls, np.uint32)
        pe, ke, _ = _arg(le, np.uint32)
        nl = len(kl)
        h = C.c_void_p()
        self._check(L.grb_collapse_f64(self.h, index.h if index is not None else None, pl, pe, nl, C.byref(h)))
        try:
            nb = L.grb_text_bytes(h)
            off = np.zeros(nl + 1, np.uint64)
            buf = np.zeros(max(nb, 1), np.uint8)
            self._check(L.grb_text_copy(h, off.ctypes.data, buf.ctypes.data))
        finally:
            L.grb_text_destroy(h)
        raw = buf.tobytes()
        return [raw[int(off[i]):int(off[i + 1])].decode() for i in range(nl)]

    def prepare_map(self, indexes, left_offsets, ls, le, ops, out, out_valid, valid_bits=False):
        """Bind every argument of grb_map_joins_f64_batch[_bits] once; calling the returned object repeats the
        query with no per-call Python work beyond the foreign call (for tight loops over a fixed batch)."""
        return PreparedMap(self, indexes, left_offsets, ls, le, ops, out, out_valid, valid_bits)

    # -- windows
    def windows_dev(self, seqlens, width, step=None, chop=False):
        sl = np.ascontiguousarray(seqlens, dtype=np.uint32)
        h = C.c_void_p()
        self._check(_capi.lib().grb_windows(self.h, sl.ctypes.data, len(sl), int(width), 0 if step is None else int(step), int(bool(chop)),
                                            C.byref(h)))
        return Windows(self, h, len(sl))

    def windows(self, seqlens, width, step=None, chop=False):
        w = self.windows_dev(seqlens, width, step, chop)
        try:
            return w.copy()
        finally:
            w.close()


    # -- merging iterators (src/merging_iterators.rs, `granges merge`)
    def merge(self, starts, ends, run_offsets=None, distance=0, scores=None, valid=None, ops=None):
        """Merge the records of a file (runs = consecutive records of one seqname; default: one run).
        -> (run_offsets_merged[nruns+1], m_starts, m_ends, first[m+1][, out[m,k], out_valid[m,k]])."""
        L = _capi.lib()
        ps, ks, _ = _arg(starts, np.uint32)
        pe, ke, _ = _arg(ends, np.uint32)
        n = len(ks)
        ro = np.ascontiguousarray([0, n] if run_offsets is None else run_offsets, dtype=np.uint64)
        nruns = len(ro) - 1
        h = C.c_void_p()
        self._check(L.grb_merge_ranges(self.h, ps, pe, ro.ctypes.data_as(_capi.u64p), nruns, int(distance), C.byref(h)))
        try:
            m = L.grb_merged_len(h)
            ms, me, first = np.empty(m, np.uint32), np.empty(m, np.uint32), np.empty(m + 1, np.uint32)
            rom = np.empty(nruns + 1, np.uint64)
            self._check(L.grb_merged_copy(h, ms.ctypes.data, me.ctypes.data, first.ctypes.data, rom.ctypes.data_as(_capi.u64p)))
            res = (rom, ms, me, first)
            if ops is not None:
                psc, ksc, _ = _arg(scores, np.float64)
                pv, kv, _ = _arg(valid, np.uint8)
                ops_a, ops_p = self._ops(ops)
                k = len(ops_a)
                out, ov = np.zeros((m, k), np.float64), np.zeros((m, k), np.uint8)
                self._check(L.grb_merged_map_f64(self.h, h, psc, pv, ops_p, k, out.ctypes.data, ov.ctypes.data))
                res = res + (out, ov)
        finally:
            L.grb_merged_destroy(h)
        return res


def map_whole_mgpu(devices, rights, left_offsets, ls, le, ops, out=None, out_valid=None):
    """`granges map` for host arrays on several GPUs of this box in one call (grb_map_joins_whole_f64_mgpu): whole seqnames are
    dealt to the GPUs in contiguous runs, every GPU writes its rows of `out` / `out_valid`.  Arguments as Context.map_whole."""
    nseq = len(rights)
    keep, ps, pe, psc, pva, ns = [], [], [], [], [], []
    any_scores = any(len(r) > 2 and r[2] is not None for r in rights)
    any_valid = any(len(r) > 3 and r[3] is not None for r in rights)
    for r in rights:
        a, ka, _ = _arg(r[0], np.uint32)
        b, kb, _ = _arg(r[1], np.uint32)
        c, kc, _ = _arg(r[2] if len(r) > 2 else None, np.float64)
        d, kd, _ = _arg(r[3] if len(r) > 3 else None, np.uint8)
        keep += [ka, kb, kc, kd]
        ps.append(a)
        pe.append(b)
        psc.append(c)
        pva.append(d)
        ns.append(len(ka))
    pl, kl, _ = _arg(ls, np.uint32)
    ple, kle, _ = _arg(le, np.uint32)
    off = np.ascontiguousarray(left_offsets, dtype=np.uint64)
    ops_a, ops_p = Context._ops(ops)
    k = len(ops_a)
    nl = len(kl)
    if out is None:
        out = np.empty((nl, k), np.float64)
    if out_valid is None:
        out_valid = np.empty((nl, k), np.uint8)
    po, pv = _out_arg(out, np.float64, (nl, k))[0], _out_arg(out_valid, np.uint8, (nl, k))[0]
    devs = np.ascontiguousarray(devices, dtype=np.int32)
    VP = C.c_void_p * nseq
    err = C.create_string_buffer(512)
    rc = _capi.lib().grb_map_joins_whole_f64_mgpu(devs.ctypes.data_as(_capi.i32p), len(devs), nseq, VP(*ps), VP(*pe), (C.c_size_t * nseq)(*ns),
                                                  VP(*psc) if any_scores else None, VP(*pva) if any_valid else None,
                                                  off.ctypes.data_as(_capi.u64p), pl, ple, ops_p, k, po, pv, err, 512)
    if rc != 0:
        raise GRangesError(rc, err.value.decode())
    return out, out_valid


class PreparedMap:
    """A grb_map_joins_f64_batch call with all arguments already marshalled."""

    def __init__(self, ctx, indexes, left_offsets, ls, le, ops, out, out_valid, valid_bits=False):
        self.ctx = ctx
        self._keep = (indexes, ls, le, out, out_valid)  # keep the buffers alive
        self._handles = Context._handles(indexes)
        self._off = np.ascontiguousarray(left_offsets, dtype=np.uint64)
        self._ops, ops_p = Context._ops(ops)
        k = len(self._ops)
        pl, _, _ = _arg(ls, np.uint32)
        pe, _, _ = _arg(le, np.uint32)
        po = _out_arg(out, np.float64)[0]
        pv = _out_arg(out_valid, np.uint32 if valid_bits else np.uint8)[0]
        self._fn = _capi.lib().grb_map_joins_f64_batch_bits if valid_bits else _capi.lib().grb_map_joins_f64_batch
        self._args = (ctx.h, self._handles, self._off.ctypes.data_as(_capi.u64p), len(indexes), pl, pe, ops_p, k, po, pv)

    def __call__(self):
        rc = self._fn(*self._args)
        if rc != 0:
            self.ctx._check(rc)


def shard_plan(left_offsets, right_counts, ls, le, nranks, max_shards=None):
    """Host-side split of one join into independent units (grb_shard_plan); returns a list of dicts."""
    off = np.ascontiguousarray(left_offsets, dtype=np.uint64)
    rc = np.ascontiguousarray(right_counts, dtype=np.uint64)
    nseq = len(off) - 1
    ls = None if ls is None else np.ascontiguousarray(ls, dtype=np.uint32)
    le = None if le is None else np.ascontiguousarray(le, dtype=np.uint32)
    cap = max_shards or (nseq + nranks + 1)
    buf = (_capi.Shard * cap)()
    n = _capi.lib().grb_shard_plan(off.ctypes.data_as(_capi.u64p), rc.ctypes.data_as(_capi.u64p), nseq,
                                   None if ls is None else ls.ctypes.data, None if le is None else le.ctypes.data, int(nranks), buf, cap)
    if n < 0:
        raise GRangesError(n, "grb_shard_plan failed")
    return [dict(seq=s.seq, rank=s.rank, left_begin=int(s.left_begin), left_end=int(s.left_end), right_lo=int(s.right_lo),
                 right_hi=int(s.right_hi)) for s in buf[:n]]
