"""Token-set Jaccard index (K3, BASELINE configs[4]): rows and queries are sets of uint32 token ids.

Runs on the same device machinery as the TF-IDF scan (text-ordered stream, chunk summaries, block-max pruning,
fused top-k) with unit weights and the Jaccard epilogue; the exact (|∩|, |∪|) integers of the returned pairs come
back too, so ``inter / union`` in float64 is bit-identical to Python's set arithmetic.  The reference has no Jaccard
path: this is an extension with unpinned parity (oracle: ``oracle.tfidf_oracle.jaccard_sets``).
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple

import numpy as np

from . import _capi


def _csr(sets: Sequence[Sequence[int]]) -> Tuple[np.ndarray, np.ndarray]:
    indptr = np.zeros(len(sets) + 1, dtype=np.int64)
    uniq = [np.unique(np.asarray(s, dtype=np.uint32)) for s in sets]
    if uniq:
        np.cumsum([len(u) for u in uniq], out=indptr[1:])
    ids = np.concatenate(uniq).astype(np.uint32) if uniq and indptr[-1] else np.zeros(0, dtype=np.uint32)
    return indptr, ids


class JaccardIndex:
    def __init__(self, vocab_size: int, device: int = 0, row_base: int = 0):
        h = C.c_void_p()
        _capi.check(_capi.load().kv_index_create(device, row_base, C.byref(h)))
        self._h, self.vocab_size = h, int(vocab_size)
        _capi.check(_capi.load().kv_index_set_mode(h, 1))

    def add_csr(self, indptr: np.ndarray, ids: np.ndarray) -> None:
        indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        tf = np.ones(len(ids), dtype=np.uint32)
        _capi.check(_capi.load().kv_index_append(self._h, indptr.ctypes.data_as(C.POINTER(C.c_int64)),
                                                 ids.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                 tf.ctypes.data_as(C.POINTER(C.c_uint32)), len(indptr) - 1))

    def add_sets(self, sets: Sequence[Sequence[int]]) -> None:
        self.add_csr(*_csr(sets))

    def finalize(self) -> None:
        _capi.check(_capi.load().kv_index_finalize(self._h, self.vocab_size))

    @property
    def n_rows(self) -> int:
        return int(_capi.load().kv_index_rows(self._h))

    def topk_csr(self, indptr: np.ndarray, ids: np.ndarray, k: int = 16):
        """(scores float32 [Q,k], rows int64 [Q,k], inter int32 [Q,k], union int32 [Q,k])."""
        lib = _capi.load()
        indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        n = len(indptr) - 1
        # ids outside the index vocabulary cannot match any row: they only count towards |q|
        inside = ids < self.vocab_size
        oov = np.zeros(n, dtype=np.float64)
        if not inside.all():
            seg = np.repeat(np.arange(n), np.diff(indptr))
            np.add.at(oov, seg[~inside], 1.0)
            keep = np.concatenate([[0], np.cumsum(inside)])
            indptr, ids = keep[indptr].astype(np.int64), ids[inside]
        tf = np.ones(len(ids), dtype=np.uint32)
        scores = np.empty((n, k), dtype=np.float32)
        rows = np.empty((n, k), dtype=np.int64)
        inter = np.empty((n, k), dtype=np.int32)
        union = np.empty((n, k), dtype=np.int32)
        p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
        _capi.check(lib.kv_topk(self._h, p(indptr, C.c_int64), p(ids, C.c_uint32), p(tf, C.c_uint32), p(oov, C.c_double), n, k,
                                p(scores, C.c_float), p(rows, C.c_int64)))
        _capi.check(lib.kv_jaccard_counts(self._h, p(indptr, C.c_int64), p(ids, C.c_uint32), p(oov, C.c_double), n, k,
                                          p(rows, C.c_int64), p(inter, C.c_int32), p(union, C.c_int32)))
        return scores, rows, inter, union

    def counts_csr(self, indptr: np.ndarray, ids: np.ndarray, rows: np.ndarray):
        """Exact (|q ∩ row|, |q ∪ row|) of given (query, GLOBAL row) pairs; -1 for rows this shard does not hold."""
        lib = _capi.load()
        indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        n, k = rows.shape
        inside = ids < self.vocab_size
        oov = np.zeros(n, dtype=np.float64)
        if not inside.all():
            seg = np.repeat(np.arange(n), np.diff(indptr))
            np.add.at(oov, seg[~inside], 1.0)
            keep = np.concatenate([[0], np.cumsum(inside)])
            indptr, ids = keep[indptr].astype(np.int64), ids[inside]
        inter = np.empty((n, k), dtype=np.int32)
        union = np.empty((n, k), dtype=np.int32)
        p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
        _capi.check(lib.kv_jaccard_counts(self._h, p(indptr, C.c_int64), p(ids, C.c_uint32), p(oov, C.c_double), n, k,
                                          p(rows, C.c_int64), p(inter, C.c_int32), p(union, C.c_int32)))
        return inter, union

    def topk_sets(self, queries: Sequence[Sequence[int]], k: int = 16):
        return self.topk_csr(*_csr(queries), k=k)

    def last_timing_ms(self):
        ms = (C.c_float * 4)()
        _capi.check(_capi.load().kv_index_last_timing(self._h, ms))
        return tuple(ms)

    def close(self) -> None:
        if self._h is not None:
            _capi.load().kv_index_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
