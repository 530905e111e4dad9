"""The oracle against every golden vector the reference produced (tests/golden/make_golden.py).

CPU only.  Pins both restatements in oracle/tfidf_oracle.py: the literal sklearn path and the
float64 closed form the CUDA kernels implement.
"""
import hashlib

import numpy as np
import pytest

from oracle import tfidf_oracle as O


def _sha(texts):
    h = hashlib.sha256()
    for t in texts:
        h.update(t.encode("utf-8"))
        h.update(b"\n")
    return h.hexdigest()


def _check_both(query, corpus, want):
    got_lit = O.score_sklearn(query, corpus)
    assert got_lit == want  # same library, same calls: bit-identical
    got_cf = O.score_closed_form(query, corpus)
    np.testing.assert_allclose(got_cf, want, rtol=1e-13, atol=1e-15)


def test_reference_unit_test_vector(golden):
    g = golden("ref_test_similarity.json")
    _check_both(g["query"], g["corpus"], g["scores"])
    assert g["scores"][0] > g["scores"][1]  # what tests/test_similarity.py:11-12 asserts
    np.testing.assert_allclose(g["scores"], [0.35505495142169013, 0.20181360410156932], rtol=1e-15)


def test_fixture54_all_queries(golden):
    g = golden("fixture54.json")
    corpus = [r["signature_text"] for r in g["records"]]
    assert len(corpus) == 54 and len(set(corpus)) == 13
    for q, want in zip(g["queries"], g["scores"]):
        _check_both(q, corpus, want)
    # top-5 rows of the first demo query and the duplicate-query score
    rows, vals = O.topk_stable(g["scores"][0], 5)
    assert rows == [20, 30, 31, 41, 51]
    assert max(g["scores"][4]) == pytest.approx(1.0, abs=1e-12)


def test_edge_cases(golden):
    g = golden("edge_cases.json")
    for q, want in zip(g["queries"], g["scores"]):
        _check_both(q, g["corpus"], want)
    assert g["empty_vocab_raises"]
    with pytest.raises(ValueError, match="empty vocabulary"):
        O.score_closed_form("a", ["b", ""])
    with pytest.raises(ValueError, match="empty vocabulary"):
        O.score_sklearn("a", ["b", ""])
    assert O.score_closed_form("anything", []) == [] and O.score_sklearn("anything", []) == []


def test_synthetic_small_and_cfg1(golden, built_lib):
    from kakveda_b200 import synth

    g = golden("synthetic_small.json")
    corpus, queries = synth.corpus(g["n"]), synth.queries(g["q"], g["n"])
    assert _sha(corpus) == g["corpus_sha256"] and _sha(queries) == g["queries_sha256"]
    st = O.CorpusStats(corpus)
    for q, want in zip(queries, g["scores"]):
        np.testing.assert_allclose(O.score_closed_form(q, corpus, st), want, rtol=1e-13, atol=1e-15)
    assert O.score_sklearn(queries[0], corpus) == g["scores"][0]

    g = golden("synthetic_cfg1.json")
    corpus, queries = synth.corpus(g["n"]), synth.queries(g["q"], g["n"])
    assert _sha(corpus) == g["corpus_sha256"] and _sha(queries) == g["queries_sha256"]
    st = O.CorpusStats(corpus)
    for i, q in enumerate(queries):
        s = O.score_closed_form(q, corpus, st)
        assert float(np.sum(s)) == pytest.approx(g["score_sums"][i], rel=1e-12)
        rows, vals = O.topk_stable(s, g["k"])
        np.testing.assert_allclose(vals, g["topk_scores"][i], rtol=1e-13)
        if i < 8:
            np.testing.assert_allclose(s, g["full_first8"][i], rtol=1e-13, atol=1e-15)
        # rows may only differ where float64 scores tie to the last ulp between the two evaluations
        for a, b in zip(rows, g["topk_rows"][i]):
            assert a == b or s[a] == pytest.approx(s[b], rel=1e-13)


def test_unpinned_class_oracles_selfconsistent():
    # dense / jaccard / hash classes have no reference implementation: parity unpinned (see module doc)
    rng = np.random.default_rng(0)
    q, c = rng.standard_normal((4, 32)).astype(np.float32), rng.standard_normal((9, 32)).astype(np.float32)
    s = O.dense_cosine(q, c)
    assert s.shape == (4, 9) and np.all(np.abs(s) <= 1 + 1e-12)
    np.testing.assert_allclose(O.dense_cosine(c, c).diagonal(), 1.0, rtol=1e-12)
    assert O.jaccard_sets([1, 2, 3], [2, 3, 4, 5]) == (2, 5)
    assert O.jaccard_sets([], []) == (0, 0)
    assert O.fingerprint64("abc") == int("ba7816bf8f01cfea", 16)
    idx, val = O.topk_rows(np.array([[0.5, 1.0, 1.0, 0.2]]), 3)
    assert idx.tolist() == [[1, 2, 0]]


def test_vectorised_closed_form(golden, built_lib):
    from kakveda_b200 import synth

    g = golden("edge_cases.json")
    np.testing.assert_allclose(O.score_matrix_closed_form(g["queries"], g["corpus"]), np.array(g["scores"]),
                               rtol=1e-12, atol=1e-15)
    g = golden("synthetic_small.json")
    corpus, queries = synth.corpus(g["n"]), synth.queries(g["q"], g["n"])
    np.testing.assert_allclose(O.score_matrix_closed_form(queries, corpus), np.array(g["scores"]), rtol=1e-12, atol=1e-15)
    g = golden("fixture54.json")
    corpus = [r["signature_text"] for r in g["records"]]
    np.testing.assert_allclose(O.score_matrix_closed_form(g["queries"], corpus), np.array(g["scores"]), rtol=1e-12, atol=1e-15)
