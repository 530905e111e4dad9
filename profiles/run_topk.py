"""Driver for ncu captures of K1b (tfidf_bound_kernel / tfidf_scan_kernel): `python profiles/run_topk.py ROWS QUERIES [K]`."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from kakveda_b200 import synth
from kakveda_b200.dist import ShardedGfkb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
q = int(sys.argv[2]) if len(sys.argv) > 2 else 16_384
k = int(sys.argv[3]) if len(sys.argv) > 3 else 16
buf, off = synth.signatures_packed(synth.CORPUS_SEED, 0, n)
sh = ShardedGfkb(device=0)
sh.build_packed(buf, off, 0)
qbuf, qoff = synth.signatures_packed(synth.QUERY_SEED, 0, q, dup_of_seed=synth.CORPUS_SEED, dup_rows=n)
qfb = sh.vocab.featurize_packed(qbuf, qoff, 0, grow=False)
sh.set_resident(qfb)
for _ in range(3):
    s, r = sh.topk_resident(k)
    torch.cuda.synchronize()
    lay = sh.index.layout()
    print("ms", sh.index.last_timing_ms(), "kernels", sh.index.last_kernel_ms(),
          {x: lay[x] for x in ("last_tiles", "last_splits", "pairs_scored", "pairs_passed_bound", "records_written")})
print("checksum", int(r.sum().item()))
