"""Driver for ncu captures of the K2 dense kernel: 1M x 768 bf16 rows, 10k queries, top-16."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from kakveda_b200 import DenseIndex

n, d, q = 1_000_000, 768, 10_000
rng = np.random.default_rng(11)
raw = rng.integers(0, 2**16, size=(n + q) * d, dtype=np.uint16)
emb = ((raw & np.uint16(0x807F)) | ((np.uint16(120) + ((raw >> np.uint16(7)) & np.uint16(7))) << np.uint16(7))).reshape(n + q, d)
dx = DenseIndex(d)
dx.add(emb[:n])
dx.finalize()
for _ in range(3):
    dx.topk(emb[n:], 16)
    print("ms", dx.last_timing())
