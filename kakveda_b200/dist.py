"""Row-sharded GFKB over the GPUs of one box.

One process per GPU (``torch.distributed``, NCCL over NVLink/NVSwitch).  Corpus rows are split
contiguously (rank r owns rows [n*r/W, n*(r+1)/W)); every rank sees the same query batch.  The
only exchanges are

* once per append epoch: an all-reduce(sum) of the per-feature document-frequency vector, because
  TF-IDF's idf and row norms use GLOBAL df and N (``allreduce_df``);
* once per query batch: ONE all-gather of the per-shard partial top-k (scores float32, rows int64),
  followed by a local merge ordered by (score desc, row asc) -- identical on every rank
  (``gather_topk`` + ``kv_merge_topk_device``).

Every rank featurises the whole corpus text so that feature ids agree without exchanging the
vocabulary (ids are deterministic, see csrc/featurizer.cpp); only its own rows go to its GPU.
The communication helpers take any torch tensors, so the world_size-2 ``gloo`` tests run them on
CPU; the compute stays in libkakveda_b200 (CUDA only).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _capi
from .similarity import FeatureBatch, GfkbIndex, Vocabulary


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    return n_rows * rank // world, n_rows * (rank + 1) // world


def allreduce_df(local_df, group=None):
    """Sum per-feature document frequencies over ranks, in place; returns the tensor."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(local_df, op=dist.ReduceOp.SUM, group=group)
    return local_df


def gather_topk(scores, rows, group=None):
    """All-gather per-shard partial top-k: [Q,k] -> ([W,Q,k] scores, [W,Q,k] rows) on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        return scores.unsqueeze(0), rows.unsqueeze(0)
    q = scores.shape[0]
    gs = torch.empty((world * q,) + tuple(scores.shape[1:]), dtype=scores.dtype, device=scores.device)
    gr = torch.empty((world * q,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    dist.all_gather_into_tensor(gs, scores.contiguous(), group=group)  # rank-major concatenation
    dist.all_gather_into_tensor(gr, rows.contiguous(), group=group)
    return gs.view((world,) + tuple(scores.shape)), gr.view((world,) + tuple(rows.shape))


def merge_on_device(device: int, gs, gr):
    """K5 on the gathered lists: ([W,Q,k],[W,Q,k]) -> ([Q,k],[Q,k])."""
    import torch

    w, q, k = gs.shape
    out_s = torch.empty((q, k), dtype=torch.float32, device=gs.device)
    out_r = torch.empty((q, k), dtype=torch.int64, device=gs.device)
    _capi.check(_capi.load().kv_merge_topk_device(device, C.c_void_p(gs.data_ptr()), C.c_void_p(gr.data_ptr()), w, q, k,
                                                  C.c_void_p(out_s.data_ptr()), C.c_void_p(out_r.data_ptr())))
    return out_s, out_r


class ShardedGfkb:
    """This rank's share of a GFKB job spread over ``world`` GPUs.

    ``mode="rows"`` (default, BASELINE configs[2]): the corpus rows are sharded, every rank scans its rows for
    ALL queries, one all-gather of partial top-k + merge.  ``mode="queries"``: every rank holds the WHOLE index
    (0.9 GB at 10M rows -- trivial next to 180 GB of HBM) and answers its contiguous slice of the query batch; the
    only exchange is the all-gather of the finished results.  Pruning thresholds are as tight as on one GPU, so
    this mode scales almost linearly; it is offered because the index is so small, not used for the headline.
    """

    def __init__(self, device: int, rank: int = 0, world: int = 1, group=None, mode: str = "rows"):
        if mode not in ("rows", "queries"):
            raise ValueError("mode must be 'rows' or 'queries'")
        self.mode = mode
        self.device, self.rank, self.world, self.group = device, rank, world, group
        self.vocab = Vocabulary()
        self.index: Optional[GfkbIndex] = None
        self.n_global = 0

    def build_packed(self, data, offsets: np.ndarray, mode: int = 0, n_threads: int = 0) -> None:
        import torch

        fb = self.vocab.featurize_packed(data, offsets, mode, grow=True, n_threads=n_threads)
        self.n_global = fb.n
        lo, hi = shard_bounds(fb.n, self.world, self.rank) if self.mode == "rows" else (0, fb.n)
        self.index = GfkbIndex(device=self.device, row_base=lo, vocab=self.vocab)
        self.index.add_features(fb, lo, hi)
        fb.close()
        if self.world > 1 and self.mode == "rows":
            df = torch.from_numpy(self.index.local_df().astype(np.int32)).to(f"cuda:{self.device}")
            allreduce_df(df, self.group)
            self.index.set_global_df(df.cpu().numpy().astype(np.uint32), self.n_global)
        self.index.finalize()

    def upload(self, qfb: FeatureBatch) -> None:
        if self.mode == "queries" and self.world > 1:
            # this rank's slice of the batch: re-pack the CSR rows [lo, hi)
            lo, hi = shard_bounds(qfb.n, self.world, self.rank)
            self._qslice = (lo, hi)
            ip = np.ascontiguousarray(qfb.indptr[lo:hi + 1])
            _capi.check(_capi.load().kv_query_upload(self.index._h, ip.ctypes.data_as(C.POINTER(C.c_int64)),
                                                     qfb.ids.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                     qfb.tf.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                     np.ascontiguousarray(qfb.oov[lo:hi]).ctypes.data_as(C.POINTER(C.c_double)),
                                                     hi - lo))
            return
        self.index.upload_queries(qfb)

    def topk_resident(self, k: int):
        """Device-only step on the uploaded batch: local scan+merge, all-gather, global merge."""
        import torch

        q = self._resident_q
        dev = f"cuda:{self.device}"
        if self.mode == "queries" and self.world > 1:
            import torch.distributed as dist

            lo, hi = self._qslice
            per = (q + self.world - 1) // self.world  # equal-sized slots for the all-gather
            s = torch.full((per, k), float("-inf"), dtype=torch.float32, device=dev)
            r = torch.full((per, k), -1, dtype=torch.int64, device=dev)
            if hi > lo:
                self.index.topk_resident(k, s.data_ptr(), r.data_ptr())
            gs = torch.empty((self.world * per, k), dtype=torch.float32, device=dev)
            gr = torch.empty((self.world * per, k), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(gs, s, group=self.group)
            dist.all_gather_into_tensor(gr, r, group=self.group)
            keep = torch.cat([torch.arange(w * per, w * per + (shard_bounds(q, self.world, w)[1] - shard_bounds(q, self.world, w)[0]),
                                           device=dev) for w in range(self.world)])
            return gs[keep], gr[keep]
        s = torch.empty((q, k), dtype=torch.float32, device=dev)
        r = torch.empty((q, k), dtype=torch.int64, device=dev)
        self.index.topk_resident(k, s.data_ptr(), r.data_ptr())
        if self.world == 1:
            return s, r
        gs, gr = gather_topk(s, r, self.group)
        return merge_on_device(self.device, gs, gr)

    def set_resident(self, qfb: FeatureBatch) -> None:
        self.upload(qfb)
        self._resident_q = qfb.n

    def topk_packed(self, data, offsets: np.ndarray, k: int, mode: int = 0):
        """End-to-end step from host text: featurise, upload, scan, exchange, merge, read back."""
        qfb = self.vocab.featurize_packed(data, offsets, mode, grow=False)
        try:
            self.set_resident(qfb)
            s, r = self.topk_resident(k)
            return s.cpu().numpy(), r.cpu().numpy()
        finally:
            qfb.close()
