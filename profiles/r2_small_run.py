"""Small pruned-path run for compute-sanitizer (racecheck / memcheck): 20k rows (625 chunks), 160 queries, k=16, plus
one K1a call and a 2-segment store match.  `compute-sanitizer --tool racecheck python profiles/r2_small_run.py`"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from kakveda_b200 import GfkbIndex, HashIndex, synth

n, q = 20000, 160
corpus, queries = synth.corpus(n), synth.queries(q, n)
ix = GfkbIndex()
ix.add_texts(corpus)
ix.finalize()
s, r = ix.topk(queries, 16)
lay = ix.layout()
assert lay["pairs_passed_bound"] > 0, "the bound kernel did not run"
full = ix.score(queries[0])
assert abs(full[r[0, 0]] - s[0, 0]) <= 1e-5 * abs(s[0, 0]) + 1e-7
hx = HashIndex()
hx.add_hashes(np.arange(1, 200001, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
print("rows", n, "queries", q, "pairs scored", lay["pairs_scored"], "launches", lay["kernel_launches"], "ok")
