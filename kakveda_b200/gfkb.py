"""Host logic of the GFKB match handler around the scan.

``match_records`` reproduces ``services/gfkb/app.py:79-102`` on top of a score vector or a
device top-k: stable sort by score descending, truncate to ``limit`` (5 in the reference),
THEN drop rows whose ``failure_type`` differs from the requested one (app.py:89-91 -- the filter
runs after the truncation, so fewer than ``limit`` matches can come back), and map each survivor
to the ``FailureMatch`` fields (app.py:92-100).  Records are plain mappings with the keys of
``CanonicalFailureRecord`` (services/shared/models.py:50-68).
"""
from __future__ import annotations

from typing import Any, List, Mapping, Optional, Sequence

MATCH_LIMIT = 5  # services/gfkb/app.py:89


def _to_match(rec: Mapping[str, Any], score: float) -> dict:
    return {
        "failure_id": rec["failure_id"],
        "version": rec["version"],
        "score": float(score),
        "failure_type": rec["failure_type"],
        "suggested_mitigation": rec.get("resolution"),
    }


def match_from_topk(records: Sequence[Mapping[str, Any]], rows: Sequence[int], scores: Sequence[float],
                    failure_type: Optional[str] = None, limit: int = MATCH_LIMIT) -> List[dict]:
    """``rows``/``scores``: one query's device top-k (k >= limit), already (score desc,row asc)."""
    out = []
    for r, s in list(zip(rows, scores))[:limit]:
        if r < 0:
            continue
        rec = records[int(r)]
        if failure_type and rec["failure_type"] != failure_type:
            continue
        out.append(_to_match(rec, s))
    return out


def match_records(engine, signature_text: str, records: Sequence[Mapping[str, Any]],
                  failure_type: Optional[str] = None, limit: int = MATCH_LIMIT) -> List[dict]:
    """The whole handler body (app.py:80-102) against an engine with the reference's ``score``."""
    if not records:
        return []
    corpus = [r["signature_text"] for r in records]
    scores = engine.score(signature_text, corpus)
    order = sorted(range(len(records)), key=lambda i: scores[i], reverse=True)[:limit]
    return match_from_topk(records, order, [scores[i] for i in order], failure_type, limit)
