// Synthetic failures.jsonl-shaped signature_text generator (test / bench support).
//
// Produces strings of the exact shape services/shared/fingerprint.py:51-66 emits --
//   "intent_tags:<sorted tags> | prompt_hint:<first 80 chars of prompt> | tools:<sorted> | env_keys:<sorted>"
// -- following the recipe of SURVEY.md section 8(d): prompt = verb + object + tail (half of the
// tails carry citation keywords so the intent tags of fingerprint.py:22-48 fire) + 0-8 words
// from a Zipf(1.1) vocabulary of 20k pseudo-words; 0-3 of 8 tools; 1-3 of 7 env keys; 30 % of
// rows are exact copies of earlier rows (GFKB appends a new version row per upsert,
// services/gfkb/app.py:132,146).  Row i is a pure function of (seed, i), so any range can be
// generated on any thread.  tests/test_synth.py checks a sample against the Python mirror of
// signature_text (and tests/golden/make_golden.py against the reference's own function).
#include "kv_internal.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

struct SplitMix {
  uint64_t s;
  explicit SplitMix(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
  }
};

const char *VERBS[10] = {"summarize", "explain", "describe", "translate", "review",
                         "classify", "rewrite", "analyze", "draft", "compare"};
const char *OBJECTS[10] = {"this paper", "the quarterly report", "the incident log", "our api docs",
                           "the customer email", "this contract", "the research notes",
                           "the release plan", "this dataset", "the meeting transcript"};
const char *TAILS[10] = {"and include citations even if none", "with references for every claim",
                         "and list the sources you used", "and add a bibliography even if not provided",
                         "and include a reference section", "in two short sentences",
                         "for a non technical reader", "as a bullet list", "without changing the meaning",
                         "and keep the original tone"};
const char *TOOLS[8] = {"search", "sql", "browser", "calculator", "code_exec", "retriever", "email", "calendar"};
const char *ENVS[7] = {"os", "region", "model", "tenant", "locale", "runtime", "gpu"};
const char *CONS = "bdfghklmnprstvwzjcx";  // 19
const char *VOW = "aeiou";

constexpr int ZIPF_V = 20000;
std::vector<double> g_cdf;
std::once_flag g_cdf_once;

void build_cdf() {
  g_cdf.resize(ZIPF_V);
  double acc = 0.0;
  for (int r = 1; r <= ZIPF_V; r++) {
    acc += std::pow((double)r, -1.1);
    g_cdf[r - 1] = acc;
  }
  for (auto &x : g_cdf) x /= acc;
}

std::string pseudo_word(int r) {  // distinct consonant-vowel words, >= 4 letters
  std::string w;
  int x = r;
  for (int i = 0; i < 2 || x > 0; i++) {
    int syl = x % 95;
    x /= 95;
    w.push_back(CONS[syl / 5]);
    w.push_back(VOW[syl % 5]);
  }
  return w;
}

bool has(const std::string &s, const char *needle) { return s.find(needle) != std::string::npos; }

// fingerprint.py:22-48 for an already-normalised prompt
std::string tags_of(const std::string &p) {
  std::vector<std::string> t;
  bool cites = has(p, "citation") || has(p, "citations") || has(p, "reference") || has(p, "references") ||
               has(p, "sources") || has(p, "bibliography");
  if (cites) t.push_back("intent:citations_required");
  if (has(p, "summarize") || has(p, "summary") || has(p, "tl;dr")) t.push_back("task:summarization");
  if (has(p, "explain") || has(p, "explanation") || has(p, "describe")) t.push_back("task:explanation");
  if (has(p, "even if not provided") || has(p, "even if none")) t.push_back("constraint:no_sources_provided");
  if (cites && has(p, "include")) t.push_back("instruction:include_references");
  std::sort(t.begin(), t.end());
  std::string out;
  for (size_t i = 0; i < t.size(); i++) {
    if (i) out.push_back(',');
    out += t[i];
  }
  return out;
}

void pick_sorted(SplitMix &rng, const char *const *pool, int pool_n, int k, std::string &out) {
  std::vector<std::string> sel;
  uint32_t used = 0;
  while ((int)sel.size() < k) {
    int j = (int)(rng.next() % (uint64_t)pool_n);
    if (used & (1u << j)) continue;
    used |= 1u << j;
    sel.emplace_back(pool[j]);
  }
  std::sort(sel.begin(), sel.end());
  for (size_t i = 0; i < sel.size(); i++) {
    if (i) out.push_back(',');
    out += sel[i];
  }
}

void fresh_row(uint64_t seed, int64_t i, std::string &out) {
  SplitMix rng(seed * 0x9E3779B97F4A7C15ULL + (uint64_t)i * 0xD1B54A32D192ED03ULL + 0x5bd1e995ULL);
  std::string prompt = VERBS[rng.next() % 10];
  prompt.push_back(' ');
  prompt += OBJECTS[rng.next() % 10];
  prompt.push_back(' ');
  prompt += TAILS[rng.next() % 10];
  int nw = (int)(rng.next() % 9);
  for (int w = 0; w < nw; w++) {
    double u = (double)(rng.next() >> 11) * (1.0 / 9007199254740992.0);
    int r = (int)(std::lower_bound(g_cdf.begin(), g_cdf.end(), u) - g_cdf.begin());
    if (r >= ZIPF_V) r = ZIPF_V - 1;
    prompt.push_back(' ');
    prompt += pseudo_word(r);
  }
  out.clear();
  out += "intent_tags:";
  out += tags_of(prompt);
  out += " | prompt_hint:";
  out.append(prompt, 0, std::min<size_t>(80, prompt.size()));
  out += " | tools:";
  pick_sorted(rng, TOOLS, 8, (int)(rng.next() % 4), out);
  out += " | env_keys:";
  pick_sorted(rng, ENVS, 7, 1 + (int)(rng.next() % 3), out);
}

void stream_row(uint64_t seed, int64_t i, uint64_t dup_of_seed, int64_t dup_rows, std::string &out) {
  SplitMix pick(seed ^ (0xA0761D6478BD642FULL * (uint64_t)(i + 1)));
  uint64_t coin = pick.next() % 100, j = pick.next();
  if (dup_of_seed != 0 && dup_rows > 0) {
    if (coin < 50) {  // a query that repeats a stored failure
      stream_row(dup_of_seed, (int64_t)(j % (uint64_t)dup_rows), 0, 0, out);
      return;
    }
  } else if (i > 0 && coin < 30) {  // a new version row of an earlier failure
    fresh_row(seed, (int64_t)(j % (uint64_t)i), out);
    return;
  }
  fresh_row(seed, i, out);
}

}  // namespace

extern "C" int kv_synth_signatures(uint64_t seed, int64_t first, int64_t count, uint64_t dup_of_seed,
                                   int64_t dup_rows, char *bytes, int64_t cap, int64_t *offsets) {
  if (count < 0 || first < 0 || !offsets || (cap > 0 && !bytes))
    return kv_fail(KV_ERR_INVALID, "kv_synth_signatures: bad arguments");
  std::call_once(g_cdf_once, build_cdf);
  int T = (int)std::thread::hardware_concurrency();
  if (T < 1) T = 1;
  if (T > 64) T = 64;
  if (count < 4096) T = 1;
  // pass 1: lengths; pass 2: write.  Rows are cheap to regenerate, so no staging buffers.
  std::vector<uint32_t> len((size_t)count);
  auto run = [&](int pass) {
    auto body = [&](int t) {
      std::string s;
      for (int64_t i = count * t / T; i < count * (t + 1) / T; i++) {
        stream_row(seed, first + i, dup_of_seed, dup_rows, s);
        if (pass == 0) len[(size_t)i] = (uint32_t)s.size();
        else memcpy(bytes + offsets[i], s.data(), s.size());
      }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; t++) th.emplace_back(body, t);
    body(0);
    for (auto &x : th) x.join();
  };
  run(0);
  offsets[0] = 0;
  for (int64_t i = 0; i < count; i++) offsets[i + 1] = offsets[i] + len[(size_t)i];
  if (offsets[count] > cap) return kv_fail(KV_ERR_NOMEM, "kv_synth_signatures: buffer too small");
  run(1);
  return KV_OK;
}
