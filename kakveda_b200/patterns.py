"""Pattern clustering on top of the all-pairs scan (BASELINE configs[3]).

The reference's pattern detector (services/pattern_detector/app.py:28-60) pulls every failure from the GFKB, keeps
those whose ``failure_type`` equals the event's, and upserts ONE named pattern when they span >= 2 apps.  This module
keeps that contract (``pattern_payload`` builds the same ``/patterns/upsert`` body: sorted unique failure_ids and
affected_apps, app.py:41-43,50-56) but can split a failure type into several patterns by similarity: every row's k
nearest other rows come from the device self-join, rows whose similarity reaches the threshold are linked, connected
components are the candidate patterns.  The grouping by similarity is an extension (the reference has none); its oracle
is a Python union-find over the float64 all-pairs matrix (tests).
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, List, Mapping, Optional, Sequence, Tuple

import numpy as np

from . import _capi


def cluster_topk(rows: np.ndarray, scores: np.ndarray, threshold: float) -> Tuple[np.ndarray, int]:
    """labels[i] = smallest row id of i's component in the graph {i ~ rows[i,j] : scores[i,j] >= threshold}."""
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    n, k = rows.shape
    labels = np.empty(n, dtype=np.int64)
    count = C.c_int64(0)
    _capi.check(_capi.load().kv_cluster_topk(n, k, rows.ctypes.data_as(C.POINTER(C.c_int64)),
                                             scores.ctypes.data_as(C.POINTER(C.c_float)), float(threshold),
                                             labels.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(count)))
    return labels, int(count.value)


def pattern_payload(name: str, records: Sequence[Mapping[str, Any]], description: Optional[str] = None) -> Dict[str, Any]:
    """The ``/patterns/upsert`` body the reference builds from a group of failures (app.py:41-43,50-56)."""
    affected = sorted(set(sum([list(r.get("affected_apps", [])) for r in records], [])))
    failure_ids = sorted(set(r.get("failure_id") for r in records if r.get("failure_id")))
    return {"name": name, "failure_ids": failure_ids, "affected_apps": affected, "description": description}


def detect_patterns(index, records: Sequence[Mapping[str, Any]], threshold: float = 0.8, k: int = 32,
                    min_apps: int = 2, failure_type: Optional[str] = None) -> List[Dict[str, Any]]:
    """Similarity-split version of pattern_detector.on_failure.

    ``index``: a finalized ``GfkbIndex`` whose row i is ``records[i]['signature_text']`` (corpus-fit mode gives a
    symmetric measure; the default mode works too).  Returns one payload per connected component that, restricted
    to ``failure_type`` (if given), spans at least ``min_apps`` apps (app.py:45-46) -- ordered by smallest row id.
    """
    scores, rows = index.selfjoin_topk(k)
    keep = np.ones(len(records), dtype=bool)
    if failure_type is not None:
        keep = np.fromiter((r.get("failure_type") == failure_type for r in records), dtype=bool, count=len(records))
        # rows of other failure types neither join nor bridge components
        bad = ~keep[np.clip(rows, 0, len(records) - 1)] | (rows < 0)
        scores = np.where(bad, -np.inf, scores).astype(np.float32)
        scores[~keep] = -np.inf
    labels, _ = cluster_topk(rows, scores, threshold)
    groups: Dict[int, List[int]] = {}
    for i, lab in enumerate(labels.tolist()):
        if keep[i]:
            groups.setdefault(lab, []).append(i)
    out = []
    for lab in sorted(groups):
        recs = [records[i] for i in groups[lab]]
        payload = pattern_payload(f"pattern-{lab:06d}", recs)
        if len(payload["affected_apps"]) >= min_apps:
            payload["rows"] = groups[lab]
            out.append(payload)
    return out
