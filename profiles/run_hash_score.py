"""Driver for ncu captures of K4 (hash scan, 64M fingerprints) and K1a (one-query float64 scan, 2M rows)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from kakveda_b200 import GfkbIndex, HashIndex, synth

rng = np.random.default_rng(11)
h = rng.integers(0, 2**63, size=64_000_000, dtype=np.uint64)
hx = HashIndex()
hx.add_hashes(h)
for _ in range(3):
    hx.match_hashes(h[rng.integers(0, len(h), size=4096)], 4)
    print("hash ms", hx.last_timing())
n = 2_000_000
buf, off = synth.signatures_packed(synth.CORPUS_SEED, 0, n)
ix = GfkbIndex()
fb = ix.vocab.featurize_packed(buf, off, 0, grow=True)
ix.add_features(fb)
fb.close()
ix.finalize()
qs = synth.queries(4, n)
for q in qs:
    ix.score(q)
    print("score ms", ix.last_score_ms())
