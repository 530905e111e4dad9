"""Driver for ncu captures / timings of K2 on device-generated embeddings.
    python profiles/run_dense_cfg.py allpairs [N] [K]     # every row's K nearest other rows (BASELINE configs[3])
    python profiles/run_dense_cfg.py scan [N] [Q] [K]     # N x 768 rows, Q queries (configs[1]/[2]-dense)"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from kakveda_b200 import DenseIndex

mode = sys.argv[1] if len(sys.argv) > 1 else "allpairs"
dd = 768
dev = torch.device("cuda:0")


def dense_rows(count, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    raw = torch.randint(0, 2**16, (count, dd), generator=g, device=dev, dtype=torch.int32)
    bits = (raw & 0x807F) | ((120 + ((raw >> 7) & 7)) << 7)
    return torch.where(bits >= 32768, bits - 65536, bits).to(torch.int16).view(torch.bfloat16).contiguous()


n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
dx = DenseIndex(dd)
for i in range(0, n, 1_000_000):
    dx.add_device(dense_rows(min(1_000_000, n - i), 100 + i // 1_000_000))
dx.finalize()
if mode == "allpairs":
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    for _ in range(2):
        s, r = dx.selfjoin_topk(k, device_out=True)
        print("allpairs ms", dx.last_timing(), "TFLOP/s", 2.0 * n * n * dd / dx.last_timing()[0] / 1e9)
else:
    q = int(sys.argv[3]) if len(sys.argv) > 3 else 100_000
    k = int(sys.argv[4]) if len(sys.argv) > 4 else 16
    qs = dense_rows(q, 999)
    for _ in range(2):
        dx.topk_device(qs, k)
        print("scan ms", dx.last_timing(), "TFLOP/s", 2.0 * n * q * dd / dx.last_timing()[0] / 1e9)
