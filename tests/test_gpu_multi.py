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
        sh = ShardedGfkb(device=rank, rank=rank, world=world, mode=mode)
        sh.build_packed(buf, off, 0, n_threads=8)
        s, r = sh.topk_packed(qbuf, qoff, k)
        np.save(Path(tmp, f"s{rank}{mode}.npy"), s)
        np.save(Path(tmp, f"r{rank}{mode}.npy"), r)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["rows", "queries"])
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
