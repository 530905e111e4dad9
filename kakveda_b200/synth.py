"""Synthetic failures.jsonl-shaped ``signature_text`` rows (test / bench support).

Thin wrapper over ``kv_synth_signatures`` (csrc/synth.cpp): row ``i`` of a stream is a pure
function of ``(seed, i)``; 30 % of corpus rows are exact copies of earlier rows (versioned
appends, services/gfkb/app.py:132,146) and -- for query streams -- half the rows are exact copies
of rows of the corpus stream.  Seeds follow SURVEY.md section 8(d).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import numpy as np

from . import _capi

CORPUS_SEED = 0xC0FFEE
QUERY_SEED = 0xFACADE


def signatures_packed(seed: int, first: int, count: int, dup_of_seed: int = 0, dup_rows: int = 0,
                      bytes_per_row: int = 288) -> Tuple[np.ndarray, np.ndarray]:
    """Rows [first, first+count) of stream ``seed`` as (uint8 buffer, int64 offsets[count+1])."""
    lib = _capi.load()
    offsets = np.zeros(count + 1, dtype=np.int64)
    cap = max(1, count * bytes_per_row)
    while True:
        buf = np.empty(cap, dtype=np.uint8)
        rc = lib.kv_synth_signatures(seed, first, count, dup_of_seed, dup_rows, buf.ctypes.data_as(C.c_char_p), cap,
                                     offsets.ctypes.data_as(C.POINTER(C.c_int64)))
        if rc == _capi.KV_ERR_NOMEM and offsets[count] > cap:
            cap = int(offsets[count])
            continue
        _capi.check(rc)
        return buf[: int(offsets[count])], offsets


def signatures(seed: int, first: int, count: int, dup_of_seed: int = 0, dup_rows: int = 0) -> List[str]:
    buf, off = signatures_packed(seed, first, count, dup_of_seed, dup_rows)
    raw = buf.tobytes()
    return [raw[off[i]:off[i + 1]].decode("ascii") for i in range(count)]


def corpus(n: int, seed: int = CORPUS_SEED) -> List[str]:
    return signatures(seed, 0, n)


def queries(q: int, corpus_rows: int, seed: int = QUERY_SEED, corpus_seed: int = CORPUS_SEED) -> List[str]:
    return signatures(seed, 0, q, dup_of_seed=corpus_seed, dup_rows=corpus_rows)
