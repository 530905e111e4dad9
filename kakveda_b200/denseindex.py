"""Dense-embedding cosine index (K2): bf16 GEMM on the tcgen05 tensor cores with a fused top-k.

Extension of the reference (which only has TF-IDF; embeddings are listed as a possible upgrade in
docs/failure-intelligence.md:43-46).  Rows and queries are float arrays rounded to bfloat16 on the way in;
the cosine is computed on those bf16 values (fp32 accumulation, fp32 norms).
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np

from . import _capi


def to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """float32 -> bfloat16 bit patterns (uint16), round-to-nearest-even."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return (((u + 0x7FFF + ((u >> 16) & 1)) >> 16) & 0xFFFF).astype(np.uint16)


class DenseIndex:
    def __init__(self, dim: int, device: int = 0, row_base: int = 0):
        h = C.c_void_p()
        _capi.check(_capi.load().kv_dense_create(device, dim, row_base, C.byref(h)))
        self._h, self.dim, self.device = h, dim, device

    def add(self, rows: np.ndarray) -> None:
        bits = rows if rows.dtype == np.uint16 else to_bf16_bits(rows)
        bits = np.ascontiguousarray(bits).reshape(-1, self.dim)
        _capi.check(_capi.load().kv_dense_append(self._h, bits.ctypes.data_as(C.POINTER(C.c_uint16)), bits.shape[0]))

    def add_device(self, rows) -> None:
        """Append rows that already live in HBM: a contiguous torch bfloat16 tensor [n, dim] on this device."""
        import torch

        assert rows.is_cuda and rows.is_contiguous() and rows.element_size() == 2 and rows.shape[-1] == self.dim
        torch.cuda.current_stream(rows.device).synchronize()  # the library copies on its own stream
        _capi.check(_capi.load().kv_dense_append_device(self._h, C.c_void_p(rows.data_ptr()), rows.shape[0]))

    def finalize(self) -> None:
        _capi.check(_capi.load().kv_dense_finalize(self._h))

    def topk_device(self, queries, k: int = 16, exclude_base: int = -1):
        """Queries and results on the device (torch): queries bfloat16 [Q, dim]; returns (float32 [Q,k], int64 [Q,k]).
        ``exclude_base >= 0``: query q never matches GLOBAL row ``exclude_base + q`` (self-join)."""
        import torch

        assert queries.is_cuda and queries.is_contiguous() and queries.element_size() == 2 and queries.shape[-1] == self.dim
        n = queries.shape[0]
        torch.cuda.current_stream(queries.device).synchronize()  # the library reads the queries on its own stream
        s = torch.empty((n, k), dtype=torch.float32, device=queries.device)
        r = torch.empty((n, k), dtype=torch.int64, device=queries.device)
        _capi.check(_capi.load().kv_dense_topk_device(self._h, C.c_void_p(queries.data_ptr()), n, k, exclude_base,
                                                      C.c_void_p(s.data_ptr()), C.c_void_p(r.data_ptr())))
        return s, r

    def selfjoin_topk(self, k: int = 32, lo: int = 0, hi: int | None = None, device_out: bool = False):
        """All-pairs (BASELINE configs[3]): for local rows [lo, hi) the k nearest OTHER rows."""
        import torch

        hi = self.n_rows if hi is None else hi
        dev = torch.device("cuda", self.device)
        s = torch.empty((hi - lo, k), dtype=torch.float32, device=dev)
        r = torch.empty((hi - lo, k), dtype=torch.int64, device=dev)
        _capi.check(_capi.load().kv_dense_selfjoin_device(self._h, lo, hi, k, C.c_void_p(s.data_ptr()), C.c_void_p(r.data_ptr())))
        return (s, r) if device_out else (s.cpu().numpy(), r.cpu().numpy())

    @property
    def n_rows(self) -> int:
        return int(_capi.load().kv_dense_rows(self._h))

    def topk(self, queries: np.ndarray, k: int = 16) -> Tuple[np.ndarray, np.ndarray]:
        bits = queries if queries.dtype == np.uint16 else to_bf16_bits(queries)
        bits = np.ascontiguousarray(bits).reshape(-1, self.dim)
        n = bits.shape[0]
        scores = np.empty((n, k), dtype=np.float32)
        rows = np.empty((n, k), dtype=np.int64)
        _capi.check(_capi.load().kv_dense_topk(self._h, bits.ctypes.data_as(C.POINTER(C.c_uint16)), n, k,
                                               scores.ctypes.data_as(C.POINTER(C.c_float)),
                                               rows.ctypes.data_as(C.POINTER(C.c_int64))))
        return scores, rows

    def last_timing(self) -> Tuple[float, int]:
        ms, sp = C.c_float(), C.c_int64()
        _capi.check(_capi.load().kv_dense_last_timing(self._h, C.byref(ms), C.byref(sp)))
        return ms.value, sp.value

    def close(self) -> None:
        if self._h is not None:
            _capi.load().kv_dense_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
