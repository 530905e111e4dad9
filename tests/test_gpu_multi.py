"""Two-rank NCCL run of the sharded GFKB (skipped on boxes with a single GPU): the df all-reduce,
the all-gather of partial top-k and the K5 merge must reproduce the single-index result."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q, k, tmp, mode):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        from kakveda_b200 import synth
        from kakveda_b200.dist import ShardedGfkb

        buf, off = synth.signatures_packed(synth.CORPUS_SEED, 0, n)
        qbuf, qoff = synth.signatures_packed(synth.QUERY_SEED, 0, q, dup_of_seed=synth.CORPUS_SEED, dup_rows=n)
        sh = ShardedGfkb(device=rank, rank=rank, world=world, mode=mode.split("-")[0], order="text" if mode.endswith("-text") else "index")
        sh.build_packed(buf, off, 0, n_threads=8)
        s, r = sh.topk_packed(qbuf, qoff, k)
        np.save(Path(tmp, f"s{rank}{mode}.npy"), s)
        np.save(Path(tmp, f"r{rank}{mode}.npy"), r)
        lay = sh.index.layout()
        sliced = int("prepare_sharded" in sh.last_e2e_ms)   # rows mode: every rank prepared only its slice of the batch
        Path(tmp, f"info{rank}{mode}.txt").write_text(f"{getattr(sh, 'n_threshold_peers', 0)} {lay['pairs_scored']} {sliced}")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["rows", "queries", "rows-text"])
def test_two_rank_sharded_topk(built_lib, tmp_path, mode):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from kakveda_b200 import GfkbIndex, synth

    n, q, k = 200_000, 1000, 16
    mp.spawn(_worker, args=(2, _free_port(), n, q, k, str(tmp_path), mode), nprocs=2, join=True)
    s0, r0 = np.load(tmp_path / f"s0{mode}.npy"), np.load(tmp_path / f"r0{mode}.npy")
    s1, r1 = np.load(tmp_path / f"s1{mode}.npy"), np.load(tmp_path / f"r1{mode}.npy")
    np.testing.assert_array_equal(r0, r1)  # every rank ends with the same merged result
    np.testing.assert_array_equal(s0, s1)
    one = GfkbIndex()
    buf, off = synth.signatures_packed(synth.CORPUS_SEED, 0, n)
    fb = one.vocab.featurize_packed(buf, off, 0, grow=True)
    one.add_features(fb)
    fb.close()
    one.finalize()
    s, r = one.topk(synth.queries(q, n), k)
    np.testing.assert_array_equal(r0, r)
    np.testing.assert_allclose(s0, s, rtol=2e-6)
    if mode.startswith("rows"):
        # the shards exchanged their pruning-threshold arrays over CUDA IPC (NVLink peer memory): one peer each
        peers = [int((tmp_path / f"info{rk}{mode}.txt").read_text().split()[0]) for rk in (0, 1)]
        assert peers == [1, 1], peers
        assert [int((tmp_path / f"info{rk}{mode}.txt").read_text().split()[2]) for rk in (0, 1)] == [1, 1]


def _dense_worker(rank, world, port, n, d, q, tmp):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        from kakveda_b200.dist import ShardedDense, shard_bounds

        g = torch.Generator().manual_seed(7)
        full = torch.randn((n + q, d), generator=g).to(torch.bfloat16)
        lo, hi = shard_bounds(n, world, rank)
        sh = ShardedDense(d, device=rank, rank=rank, world=world)
        sh.build(full[lo:hi].to(f"cuda:{rank}"), n)
        s, r = sh.topk(full[n:].to(f"cuda:{rank}"), 16)
        sa, ra = sh.allpairs_topk(32, block=700)
        np.save(Path(tmp, f"ds{rank}.npy"), s.cpu().numpy()); np.save(Path(tmp, f"dr{rank}.npy"), r.cpu().numpy())
        np.save(Path(tmp, f"as{rank}.npy"), sa.cpu().numpy()); np.save(Path(tmp, f"ar{rank}.npy"), ra.cpu().numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_dense_and_allpairs(built_lib, tmp_path):
    """BASELINE configs[2]/[3] plumbing for K2: row shards + all-gather + K5 == one index; all-pairs with self excluded."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from kakveda_b200 import DenseIndex

    n, d, q = 3001, 128, 300
    mp.spawn(_dense_worker, args=(2, _free_port(), n, d, q, str(tmp_path)), nprocs=2, join=True)
    g = torch.Generator().manual_seed(7)
    full = torch.randn((n + q, d), generator=g).to(torch.bfloat16)
    one = DenseIndex(d)
    one.add_device(full[:n].to("cuda:0").contiguous())
    one.finalize()
    s, r = one.topk_device(full[n:].to("cuda:0").contiguous(), 16)
    sa, ra = one.selfjoin_topk(32)
    for rank in (0, 1):
        np.testing.assert_array_equal(np.load(tmp_path / f"dr{rank}.npy"), r.cpu().numpy())
        np.testing.assert_array_equal(np.load(tmp_path / f"ds{rank}.npy"), s.cpu().numpy())
        np.testing.assert_array_equal(np.load(tmp_path / f"ar{rank}.npy"), ra)
        np.testing.assert_array_equal(np.load(tmp_path / f"as{rank}.npy"), sa)


def _jaccard_sets(n, q, seed=3):
    rng = np.random.default_rng(seed)
    draws = np.minimum(rng.zipf(1.3, size=(n + q, 24)) - 1, 4095).astype(np.uint32)
    draws.sort(axis=1)
    keep = np.ones(draws.shape, dtype=bool)
    keep[:, 1:] = draws[:, 1:] != draws[:, :-1]
    indptr = np.concatenate([[0], np.cumsum(keep.sum(axis=1))]).astype(np.int64)
    return indptr, draws[keep]


def _jaccard_worker(rank, world, port, n, q, tmp):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        from kakveda_b200.dist import ShardedJaccard

        indptr, ids = _jaccard_sets(n, q)
        sh = ShardedJaccard(4096, device=rank, rank=rank, world=world)
        sh.build_csr(indptr[: n + 1], ids[: indptr[n]])
        out = sh.topk_csr(indptr[n:] - indptr[n], ids[indptr[n]:], 16)
        for name, a in zip("srio", out):
            np.save(Path(tmp, f"j{name}{rank}.npy"), a)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_jaccard(built_lib, tmp_path):
    """BASELINE configs[4] plumbing: row-sharded token sets, exact (|intersection|, |union|) integers after the merge."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from kakveda_b200 import JaccardIndex

    n, q = 20000, 200
    mp.spawn(_jaccard_worker, args=(2, _free_port(), n, q, str(tmp_path)), nprocs=2, join=True)
    indptr, ids = _jaccard_sets(n, q)
    one = JaccardIndex(4096)
    one.add_csr(indptr[: n + 1], ids[: indptr[n]])
    one.finalize()
    s, r, inter, union = one.topk_csr(indptr[n:] - indptr[n], ids[indptr[n]:], 16)
    for rank in (0, 1):
        np.testing.assert_array_equal(np.load(tmp_path / f"jr{rank}.npy"), r)
        np.testing.assert_array_equal(np.load(tmp_path / f"js{rank}.npy"), s)
        np.testing.assert_array_equal(np.load(tmp_path / f"ji{rank}.npy"), inter)
        np.testing.assert_array_equal(np.load(tmp_path / f"jo{rank}.npy"), union)
    # and the integers are Python-set exact
    for qi in range(0, q, 17):
        qs = set(ids[indptr[n + qi]: indptr[n + qi + 1]].tolist())
        for j in range(16):
            rs = set(ids[indptr[r[qi, j]]: indptr[r[qi, j] + 1]].tolist())
            assert (inter[qi, j], union[qi, j]) == (len(qs & rs), len(qs | rs))
