"""GPU parity of the widened surface (SURVEY.md section 8(f) + BASELINE configs[3]): the resident GFKB store against
the match / upsert stream the REFERENCE produced (tests/golden/service_upsert.json), the statistics-only finalize,
K6 re-scoring, the self-join (exclusions, corpus-fit TF-IDF) against sklearn goldens, pattern clustering, batched
warnings, and the dense K2 extensions (k = 32, self exclusion, device queries)."""
import numpy as np
import pytest

from oracle import tfidf_oracle as O
from test_gpu_parity import check_topk, RTOL64

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib(built_lib):
    from kakveda_b200 import _capi

    assert _capi.load().kv_device_count() > 0, "GPU tests need a CUDA device"
    return _capi.load()


def _replay(golden, tail_limit, tmp_path):
    from kakveda_b200 import GfkbStore

    g = golden("service_upsert.json")
    st = GfkbStore(path=tmp_path / f"f{tail_limit}.jsonl", tail_limit=tail_limit)
    seen = []
    for step in g["steps"]:
        out = st.upsert(step["upsert"])
        assert out["created"] == step["created"]
        for m in step.get("matches", []):
            got = st.match(m["signature_text"], m["failure_type"])
            want = m["matches"]
            assert [(x["failure_id"], x["version"], x["failure_type"], x["suggested_mitigation"]) for x in got] == \
                   [(x["failure_id"], x["version"], x["failure_type"], x["suggested_mitigation"]) for x in want], m
            np.testing.assert_allclose([x["score"] for x in got], [x["score"] for x in want], rtol=RTOL64)
            ex = st.match_exact(m["signature_text"], m["failure_type"])
            assert [(x["failure_id"], x["version"]) for x in ex] == [(x["failure_id"], x["version"]) for x in want]
            np.testing.assert_allclose([x["score"] for x in ex], [x["score"] for x in want], rtol=RTOL64)
            seen.append([(x["failure_id"], x["version"], x["score"]) for x in got])
    return st, seen


def test_store_replays_reference_upserts_and_matches(lib, golden, tmp_path):
    """main + tail segments with statistics-only refreshes of the main segment == the reference handlers."""
    st, seen = _replay(golden, 40, tmp_path)
    assert st.stats["stat_refreshes"] > 0 and st.stats["compactions"] > 0
    # always rebuilding one index (tail_limit = 0) must give the same answers: same ids, float64 scores to ~1 ulp
    st0, seen0 = _replay(golden, 0, tmp_path)
    assert st0.stats["stat_refreshes"] == 0
    for a, b in zip(seen, seen0):
        assert [(x[0], x[1]) for x in a] == [(x[0], x[1]) for x in b]
        # a score may come from K6 (candidate path) in one store and from K1a (exact path) in the other: both float64, they
        # agree to 1e-13 (test_rescore_is_float64_and_ties_exactly); one decade of slack on top
        np.testing.assert_allclose([x[2] for x in a], [x[2] for x in b], rtol=1e-12)
    # a store opened on the written file serves the same matches (cold start: one full build)
    from kakveda_b200 import GfkbStore
    st2 = GfkbStore(path=st.path)
    assert len(st2.records) == len(st.records)
    q = st.records[3]["signature_text"]
    assert [(m["failure_id"], m["version"]) for m in st2.match(q)] == [(m["failure_id"], m["version"]) for m in st.match(q)]


def test_store_batch_equals_single_and_fixture(lib, golden):
    from kakveda_b200 import GfkbStore

    g = golden("fixture54.json")
    st = GfkbStore()
    st._reset([dict(r) for r in g["records"]])
    cases = g["match"]
    got = st.match_batch([c["signature_text"] for c in cases], [c["failure_type"] for c in cases])
    for c, ms in zip(cases, got):
        assert [(m["failure_id"], m["version"]) for m in ms] == [(m["failure_id"], m["version"]) for m in c["matches"]]
        np.testing.assert_allclose([m["score"] for m in ms], [m["score"] for m in c["matches"]], rtol=RTOL64)
    # batched warnings (services/warning_policy/app.py:19-72): threshold, message text, references
    reqs = [{"app_id": "a", "prompt": "Summarize this paper and include citations even if none", "tools": [], "env": {"os": "linux"}},
            {"app_id": "a", "prompt": "completely unrelated gardening advice", "tools": ["x"], "env": {}}]
    pats = [{"name": "Citation hallucination without sources", "pattern_id": "P-1"}]
    w = st.warn_batch(reqs, threshold=0.5, default_action="block", patterns=pats)
    assert w[0]["action"] == "block" and w[0]["references"] and w[0]["pattern_id"] == "P-1"
    best = w[0]["references"][0]
    assert w[0]["message"] == (f"This execution matches past failure type {best['failure_type']} (failure_id={best['failure_id']}, "
                               f"similarity={best['score']:.2f}). Suggested mitigation: {best['suggested_mitigation'] or 'n/a'}")
    assert w[0]["confidence"] == pytest.approx(0.69528258, rel=1e-6)   # the fixture's top score for this prompt
    assert w[1]["action"] == "warn" and w[1]["references"] == [] and w[1]["message"] == "No high-similarity match found in GFKB."
    assert st.warn_batch(reqs[:1], threshold=0.9)[0]["references"] == []


def test_stats_only_finalize_equals_full_rebuild(lib, monkeypatch):
    from kakveda_b200 import GfkbIndex, Vocabulary, synth

    n, extra = 30000, 700
    corpus = synth.corpus(n + extra)
    queries = synth.queries(200, n + extra)

    def build(force_full):
        if force_full:
            monkeypatch.setenv("KAKVEDA_B200_FULL_FINALIZE", "1")
        else:
            monkeypatch.delenv("KAKVEDA_B200_FULL_FINALIZE", raising=False)
        vocab = Vocabulary()
        main = GfkbIndex(vocab=vocab)
        main.add_texts(corpus[:n])
        main.finalize()
        tail = GfkbIndex(row_base=n, vocab=vocab)
        tail.add_texts(corpus[n:])
        df = np.zeros(len(vocab), dtype=np.int64)
        d0 = main.local_df()
        df[:len(d0)] += d0
        df += tail.local_df()
        for ix in (main, tail):
            ix.set_global_df(df.astype(np.uint32), n + extra)
            ix.finalize()
        return main, tail

    main, tail = build(False)
    assert main.last_finalize_kind == 2 and tail.last_finalize_kind == 1
    fmain, ftail = build(True)
    assert fmain.last_finalize_kind == 1
    for a, b in ((main, fmain), (tail, ftail)):
        s1, r1 = a.topk(queries, 16)
        s2, r2 = b.topk(queries, 16)
        np.testing.assert_array_equal(r1, r2)
        np.testing.assert_array_equal(s1, s2)
        np.testing.assert_array_equal(a.score(queries[0]), b.score(queries[0]))
    # and both equal ONE index over all rows (the oracle of the segment scheme)
    oracle = O.score_matrix_closed_form(queries[:40], corpus)
    np.testing.assert_allclose(np.concatenate([main.score(queries[1]), tail.score(queries[1])]), oracle[1], rtol=RTOL64, atol=1e-15)
    s, r = main.topk(queries[:40], 16)
    check_topk(s, r, oracle[:, :n], 16)


def test_rescore_is_float64_and_ties_exactly(lib):
    from kakveda_b200 import GfkbIndex, synth

    n, q, k = 20000, 64, 16
    corpus, queries = synth.corpus(n), synth.queries(q, n)
    corpus[100] = corpus[7]
    corpus[19000] = corpus[7]
    queries[0] = corpus[7]
    ix = GfkbIndex()
    ix.add_texts(corpus)
    ix.finalize()
    fb = ix.vocab.featurize(queries, grow=False)
    s32, rows = ix.topk_features(fb, k)
    f64 = ix.rescore(fb, rows)
    fb.close()
    oracle = O.score_matrix_closed_form(queries, corpus)
    np.testing.assert_allclose(f64, np.take_along_axis(oracle, rows, axis=1), rtol=1e-12)
    np.testing.assert_allclose(f64, s32, rtol=1e-5)
    dup = [j for j in range(k) if rows[0, j] in (7, 100, 19000)]
    assert len(dup) == 3 and len({f64[0, j] for j in dup}) == 1            # identical rows: identical bits
    full = ix.score(queries[0])
    np.testing.assert_allclose(f64[0], full[rows[0]], rtol=1e-13)  # K1a: exact fixed-point sums (2^-50), K6: ordered float64 sums


def test_selfjoin_corpus_fit_matches_sklearn_golden(lib, golden):
    from kakveda_b200 import GfkbIndex, patterns, synth

    g = golden("corpus_fit.json")
    n, k = g["n"], g["k"]
    corpus = synth.corpus(n)
    ix = GfkbIndex()
    ix.add_texts(corpus)
    ix.set_mode(2)
    ix.finalize()
    S = O.corpus_fit_scores(corpus, corpus)
    np.testing.assert_allclose(S.sum(axis=1), g["row_sums"], rtol=1e-12)   # the oracle is the pinned sklearn path
    scores, rows = ix.selfjoin_topk(k)
    assert not np.any(rows == np.arange(n)[:, None])                        # a row never matches itself
    Sx = S.copy()
    np.fill_diagonal(Sx, -np.inf)
    check_topk(scores, rows, Sx, k)
    np.testing.assert_allclose(scores, np.array(g["allpairs_scores"]), rtol=1e-5, atol=1e-7)
    # corpus-fit query scoring (fixed idf, out-of-vocabulary query features ignored)
    qs = synth.queries(6, n)
    for q, want in zip(qs, g["query_scores"]):
        np.testing.assert_allclose(ix.score(q), want, rtol=RTOL64, atol=1e-15)
    # clustering on the device top-k == union-find on the float64 matrix restricted to the same lists
    for thr in (0.6, 0.8, 0.95):
        labels, count = patterns.cluster_topk(rows, scores, thr)
        want = O.components(n, rows, scores, thr)
        np.testing.assert_array_equal(labels, want)
        assert count == len(set(want))


def test_selfjoin_default_mode_and_exclusions(lib):
    from kakveda_b200 import GfkbIndex, synth

    n, k = 5000, 8
    corpus = synth.corpus(n)
    ix = GfkbIndex()
    ix.add_texts(corpus)
    ix.finalize()
    s_self, r_self = ix.selfjoin_topk(k, 1000, 1600)
    s_all, r_all = ix.topk(corpus[1000:1600], k + 1)
    for i in range(600):
        keep = [j for j in range(k + 1) if r_all[i, j] != 1000 + i][:k]
        assert r_self[i].tolist() == r_all[i, keep].tolist()
        np.testing.assert_array_equal(s_self[i], s_all[i, keep])
    # explicit exclusions of arbitrary rows, and clearing them
    fb = ix.vocab.featurize(corpus[:50], grow=False)
    ix.upload_queries(fb)
    ix.set_exclusions(np.arange(50) + 1)            # forbid row q+1 instead of row q
    s, r = ix.topk_resident_host(50, k)
    assert not np.any(r == (np.arange(50) + 1)[:, None])
    assert all(corpus[r[i, 0]] == corpus[i] and r[i, 0] <= i for i in range(50))   # best hit: the row itself or an earlier duplicate
    ix.set_exclusions(None)
    s2, r2 = ix.topk_resident_host(50, k)
    s3, r3 = ix.topk_features(fb, k)
    fb.close()
    np.testing.assert_array_equal(r2, r3)
    # null queries skip the excluded row too
    fbn = ix.vocab.featurize(["", "zzzz qqqq"], grow=False)
    ix.upload_queries(fbn)
    ix.set_exclusions(np.array([0, 2]))
    s, r = ix.topk_resident_host(2, 4)
    fbn.close()
    assert r.tolist() == [[1, 2, 3, 4], [0, 1, 3, 4]]


def test_detect_patterns_splits_by_similarity(lib):
    from kakveda_b200 import GfkbIndex, patterns, synth

    n = 600
    corpus = synth.corpus(n)
    records = [{"failure_id": f"F-{i + 1:04d}", "failure_type": "HALLUCINATION_CITATION" if i % 3 else "OTHER",
                "affected_apps": [f"app-{i % 5}"], "signature_text": t} for i, t in enumerate(corpus)]
    ix = GfkbIndex()
    ix.add_texts(corpus)
    ix.set_mode(2)
    ix.finalize()
    # pick a threshold that no candidate score comes close to, so float32 vs float64 cannot disagree about an edge
    S = O.corpus_fit_scores(corpus, corpus)
    s_dev, r_dev = ix.selfjoin_topk(16)
    f64 = S[np.arange(n)[:, None], np.clip(r_dev, 0, n - 1)]
    allv = np.unique(f64.ravel())
    i = int(np.searchsorted(allv, 0.75))
    while allv[i] - allv[i - 1] < 1e-3:
        i += 1
    thr = float((allv[i] + allv[i - 1]) / 2)
    out = patterns.detect_patterns(ix, records, threshold=thr, k=16, failure_type="HALLUCINATION_CITATION")
    keep = np.array([r["failure_type"] == "HALLUCINATION_CITATION" for r in records])
    vals = np.where(keep[np.clip(r_dev, 0, n - 1)] & keep[:, None] & (r_dev >= 0), f64, -np.inf)
    labels = O.components(n, r_dev, vals, thr)
    groups = {}
    for i, lab in enumerate(labels):
        if keep[i]:
            groups.setdefault(lab, []).append(i)
    want = [g for _, g in sorted(groups.items()) if len({records[i]["affected_apps"][0] for i in g}) >= 2]
    assert [p["rows"] for p in out] == want
    for p in out:
        assert p["failure_ids"] == sorted(records[i]["failure_id"] for i in p["rows"])


def test_dense_k32_exclusion_and_device_queries(lib):
    import torch
    from kakveda_b200 import DenseIndex

    rng = np.random.default_rng(3)
    n, d, q = 3000, 128, 200
    base = rng.standard_normal((40, d)).astype(np.float32)
    rows = (base[rng.integers(0, 40, n)] + 0.3 * rng.standard_normal((n, d))).astype(np.float32)
    rows[50] = rows[10]                                   # duplicates tie -> lower row first
    dx = DenseIndex(d)
    dx.add(rows)
    dx.finalize()
    S = O.dense_cosine(rows, rows)
    # k = 32 from host queries
    s, r = dx.topk(rows[:q], 32)
    check_topk(s, r, S[:q], 32, rtol=2e-5)
    # the same through device tensors
    tq = torch.from_numpy(O.bf16_round(rows[:q])).to("cuda").to(torch.bfloat16).contiguous()
    sd, rd = dx.topk_device(tq, 32)
    np.testing.assert_array_equal(rd.cpu().numpy(), r)
    np.testing.assert_array_equal(sd.cpu().numpy(), s)
    # self-join: every row's nearest OTHER rows
    ss, rs = dx.selfjoin_topk(32)
    assert not np.any(rs == np.arange(n)[:, None])
    Sx = S.copy()
    np.fill_diagonal(Sx, -np.inf)
    check_topk(ss, rs, Sx, 32, rtol=2e-5)
    assert rs[10, 0] == 50 and rs[50, 0] == 10
    # a device-appended copy of the rows behaves identically
    dy = DenseIndex(d, row_base=1000)
    dy.add_device(torch.from_numpy(O.bf16_round(rows)).to("cuda").to(torch.bfloat16).contiguous())
    dy.finalize()
    s2, r2 = dy.topk_device(tq, 16, exclude_base=1000)
    for i in range(q):
        assert 1000 + i not in r2[i].tolist()
    np.testing.assert_array_equal(r2.cpu().numpy()[:, :8] - 1000, rs[:q, :8])



def test_persisted_layout_gives_identical_results(lib, tmp_path):
    """an index whose scan layout was restored from the layout file (no host sort / block build:
    kv_index_last_finalize_kind == 2) returns bit-identical top-k and float64 scores; a file built for other rows is
    refused; GfkbStore restores it on a cold start next to its sidecar."""
    from kakveda_b200 import GfkbIndex, synth

    n, q, k = 40_000, 500, 16
    corpus, queries = synth.corpus(n), synth.queries(q, n)
    a = GfkbIndex()
    a.add_texts(corpus)
    a.finalize()
    assert a.last_finalize_kind == 1
    s1, r1 = a.topk(queries, k)
    f1 = a.score(queries[3])
    path = tmp_path / "gfkb.layout"
    a.save_layout(path)
    b = GfkbIndex()
    b.add_texts(corpus)
    assert b.load_layout(path)
    b.finalize()
    assert b.last_finalize_kind == 2          # statistics only: nothing was sorted or rebuilt
    s2, r2 = b.topk(queries, k)
    np.testing.assert_array_equal(r1, r2)
    np.testing.assert_array_equal(s1, s2)
    np.testing.assert_array_equal(f1, b.score(queries[3]))
    c = GfkbIndex()
    c.add_texts(corpus[:-1] + ["intent_tags: | prompt_hint:some other row | tools: | env_keys:os"])
    assert not c.load_layout(path)            # other rows: refused, the usual finalize follows
    c.finalize()
    assert c.last_finalize_kind == 1
    assert not GfkbIndex().load_layout(tmp_path / "missing.layout")


def test_warn_batch_equals_reference_warn_handler(lib, golden):
    """services/warning_policy/app.py:19-72 run UNMODIFIED against the reference GFKB (tests/golden/make_golden_warn.py):
    action, confidence, pattern id, references and the message text of every response are reproduced by
    GfkbStore.warn_batch under the three recorded policy configurations."""
    from kakveda_b200 import GfkbStore

    g = golden("service_warn.json")
    st = GfkbStore()
    st._reset([dict(r) for r in g["failures"]])
    for pol in g["policies"]:
        got = st.warn_batch(g["requests"], threshold=pol["threshold"], default_action=pol["default_action"], patterns=[g["pattern"]])
        assert len(got) == len(pol["responses"])
        for a, b in zip(got, pol["responses"]):
            assert a["action"] == b["action"] and a["pattern_id"] == b["pattern_id"]
            np.testing.assert_allclose(a["confidence"], b["confidence"], rtol=RTOL64, atol=1e-15)
            assert len(a["references"]) == len(b["references"])
            for ra, rb in zip(a["references"], b["references"]):
                assert (ra["failure_id"], ra["version"], ra["failure_type"], ra["suggested_mitigation"]) == \
                       (rb["failure_id"], rb["version"], rb["failure_type"], rb["suggested_mitigation"])
                np.testing.assert_allclose(ra["score"], rb["score"], rtol=RTOL64, atol=1e-15)
            if b["references"]:
                # the text embeds the score with two decimals; everything else must match literally
                assert a["message"] == b["message"]
            else:
                assert a["message"] == b["message"] == "No high-similarity match found in GFKB."
