"""Mnemonic counts of the hot kernels from `cuobjdump -sass` of the built library (evidence that the tensor-core, TMA
and mbarrier instructions are in the shipped code and that shared-memory accesses are LDS/STS, not generic LD/ST).
Run:  python profiles/r2_sass_counts.py > profiles/r2_sass_hot_kernels.txt"""
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "kakveda_b200" / "lib" / "libkakveda_b200.so"
HOT = ("dense_topk_kernel", "hash_scan_kernel", "tfidf_select_kernel", "tfidf_bound_kernel", "tfidf_scan_kernel",
       "jaccard_scan_kernel", "tfidf_score_kernel", "prep_queries_kernel", "merge_topk_kernel")
SHOW = ("UTCHMMA", "UTCBAR", "UTMALDG", "UBLKCP", "LDTM", "SYNCS", "ATOMS", "LDS", "STS", "LDG", "STG", "ATOMG", "REDG", "RED", "SHFL",
        "VOTE", "R2P", "BAR", "LD", "ST", "ATOM", "FADD", "IADD3", "IMAD", "MUFU")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    print("cuobjdump -sass kakveda_b200/lib/libkakveda_b200.so (sm_100a), mnemonic counts of the hot kernels")
    print("(UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA tile), UBLKCP = cp.async.bulk (TMA 1-D), "
          "SYNCS = mbarrier, LD/ST/ATOM without S/G = generic address space)\n")
    name, counts = None, Counter()

    def flush():
        if name and any(h in name for h in HOT):
            print(name)
            print(f"  instructions: {sum(counts.values())}")
            print("  " + "  ".join(f"{m}={counts[m]}" for m in SHOW if counts[m]) + "\n")

    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            flush()
            name, counts = m.group(1), Counter()
            continue
        m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            counts[m.group(1)] += 1
    flush()


if __name__ == "__main__":
    sys.exit(main())
