"""Host-side producer of the strings the GFKB scan matches.

Mirrors the behaviour of the reference's ``services/shared/fingerprint.py``:

* ``normalize_prompt``   -- fingerprint.py:16-19 (strip, lower, collapse whitespace runs)
* ``intent_tags``        -- fingerprint.py:22-48 (coarse keyword tags, sorted, unique)
* ``signature_text``     -- fingerprint.py:51-66 (``intent_tags:.. | prompt_hint:<80 chars> | tools:.. | env_keys:..``)
* ``fingerprint``        -- fingerprint.py:69-71 (first 16 hex digits of sha256(signature_text))

These run once per request / stored failure on the CPU and stay there: they define the
row shape (<= ~60 word 1,2-gram features per row) the device index is laid out for.
"""
from __future__ import annotations

import hashlib
import re
from typing import Any, Iterable, List, Mapping

_WS = re.compile(r"\s+")

# (tag, keywords any-of) -- evaluated on the normalised prompt
_CITATION_WORDS = ("citation", "citations", "reference", "references", "sources", "bibliography")
_TASK_RULES = (
    ("task:summarization", ("summarize", "summary", "tl;dr")),
    ("task:explanation", ("explain", "explanation", "describe")),
)
_NO_SOURCE_PHRASES = ("even if not provided", "even if none")
PROMPT_HINT_CHARS = 80


def normalize_prompt(prompt: str) -> str:
    return _WS.sub(" ", prompt.strip().lower())


def intent_tags(prompt: str) -> List[str]:
    p = normalize_prompt(prompt)
    found = set()
    cites = any(w in p for w in _CITATION_WORDS)
    if cites:
        found.add("intent:citations_required")
    for tag, words in _TASK_RULES:
        if any(w in p for w in words):
            found.add(tag)
    if any(ph in p for ph in _NO_SOURCE_PHRASES):
        found.add("constraint:no_sources_provided")
    if cites and "include" in p:
        found.add("instruction:include_references")
    return sorted(found)


def signature_text(prompt: str, tools: Iterable[str], env: Mapping[str, Any]) -> str:
    hint = normalize_prompt(prompt)[:PROMPT_HINT_CHARS]
    fields = (
        ("intent_tags", ",".join(intent_tags(prompt))),
        ("prompt_hint", hint),
        ("tools", ",".join(sorted(set(tools)))),
        ("env_keys", ",".join(sorted(env.keys()))),
    )
    return " | ".join(f"{k}:{v}" for k, v in fields)


def fingerprint_text(sig: str) -> str:
    """64-bit hash fingerprint (hex) of an already-built signature_text."""
    return hashlib.sha256(sig.encode("utf-8")).hexdigest()[:16]


def fingerprint(prompt: str, tools: Iterable[str], env: Mapping[str, Any]) -> str:
    return fingerprint_text(signature_text(prompt, tools, env))


def fingerprint_u64(sig: str) -> int:
    """The same 64 bits as an integer -- what the device hash index (K4) stores per row."""
    return int(fingerprint_text(sig), 16)
