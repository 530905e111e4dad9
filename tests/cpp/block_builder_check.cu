// CPU check (no GPU needed) of the host-side scan-layout construction in kakveda_b200/csrc/block_builder.cuh:
//   1. decoding the column blocks of a random corpus gives back every row's (feature, tf) set exactly (universal
//      features excluded), for full and partial last chunks, empty rows and term frequencies beyond the 5-bit field;
//   2. block invariants: rare entries first, both parts sorted by (feature, tf), W_ALL exactly when the mask covers
//      every valid row, 16-byte block alignment, the dense matrix Uf holds the largest tf of each frequent feature;
//   3. the multi-threaded build equals the single-threaded one byte for byte;
//   4. the fixed-point row sums the scan kernels form (integer weights, any order) agree with float64 to 1e-9.
// Exit code 0 = all good.  Run by tests/test_host.py::test_block_builder_roundtrip.
#include "../../kakveda_b200/csrc/block_builder.cuh"

#include <cstdio>
#include <map>
#include <random>
#include <set>

using namespace kvh;

static int fails = 0;
#define CHECK(c, ...) do { if (!(c)) { if (fails < 20) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } fails++; } } while (0)

struct Corpus {
  std::vector<int64_t> indptr{0};
  std::vector<uint32_t> ids;
  std::vector<uint16_t> tf;
  std::vector<uint8_t> univ;
  std::vector<uint32_t> tfmax;
  std::vector<int> perm;
  int64_t n = 0, V = 0;
};

static Corpus make(int64_t n, int64_t V, int n_univ, unsigned seed, bool big_tf) {
  Corpus c;
  c.n = n; c.V = V;
  std::mt19937 rng(seed);
  c.univ.assign((size_t)V, 0);
  for (int u = 0; u < n_univ; u++) c.univ[(size_t)u] = 1;
  c.tfmax.assign((size_t)V, 0);
  for (int64_t r = 0; r < n; r++) {
    std::map<uint32_t, uint16_t> row;
    for (int u = 0; u < n_univ; u++) row[(uint32_t)u] = 1;
    int len = (r % 17 == 5) ? 0 : 3 + (int)(rng() % 30);
    for (int j = 0; j < len; j++) {
      // a Zipf-like mix: few very common features, many rare ones
      uint32_t f = (rng() % 3 == 0) ? (uint32_t)(n_univ + rng() % 40) : (uint32_t)(n_univ + rng() % (V - n_univ));
      uint16_t t = 1;
      if (rng() % 11 == 0) t = (uint16_t)(1 + rng() % 5);
      if (big_tf && rng() % 97 == 0) t = (uint16_t)(31 + rng() % 3000);
      row[f] = t;
    }
    for (auto &kv : row) {
      c.ids.push_back(kv.first);
      c.tf.push_back(kv.second);
      c.tfmax[kv.first] = std::max<uint32_t>(c.tfmax[kv.first], kv.second);
    }
    c.indptr.push_back((int64_t)c.ids.size());
  }
  c.perm.resize((size_t)n);
  for (int64_t i = 0; i < n; i++) c.perm[(size_t)i] = (int)i;
  std::shuffle(c.perm.begin(), c.perm.end(), rng);
  return c;
}

static std::vector<uint32_t> flat(const BlockLayout &L) {
  std::vector<uint32_t> out((size_t)L.total_words);
  for (size_t t = 0; t < L.parts.size(); t++)
    if (!L.parts[t].empty()) memcpy(out.data() + L.part_off[t], L.parts[t].data(), L.parts[t].size() * 4);
  return out;
}

static uint32_t ovf_get(const BlockLayout &L, int64_t c, uint32_t e) {
  auto key = ((unsigned long long)c << 32) | e;
  auto it = std::lower_bound(L.ovf.begin(), L.ovf.end(), std::make_pair(key, 0u));
  if (it == L.ovf.end() || it->first != key) return 0;
  return it->second;
}

static void check_case(int64_t n, int64_t V, int n_univ, unsigned seed, bool big_tf) {
  Corpus c = make(n, V, n_univ, seed, big_tf);
  BlockLayout L1, L4;
  build_blocks(c.indptr.data(), c.ids.data(), c.tf.data(), c.perm.data(), n, V, c.univ.data(), c.tfmax.data(), 1, L1);
  build_blocks(c.indptr.data(), c.ids.data(), c.tf.data(), c.perm.data(), n, V, c.univ.data(), c.tfmax.data(), 4, L4);
  std::vector<uint32_t> b1 = flat(L1), b4 = flat(L4);
  CHECK(b1 == b4, "threaded build differs (n=%lld)", (long long)n);
  CHECK(L1.ovf == L4.ovf && L1.fslot == L4.fslot, "threaded build: overflow table / columns differ");
  CHECK(memcmp(L1.binfo.data(), L4.binfo.data(), L1.binfo.size() * sizeof(BlockInfo)) == 0, "threaded build: directory differs");
  CHECK(memcmp(L1.Uf.data(), L4.Uf.data(), L1.Uf.size() * sizeof(__half)) == 0, "threaded build: dense matrix differs");
  CHECK(L1.fslot2 == L4.fslot2 && L1.Ubt == L4.Ubt, "threaded build: second-class bitmaps differ");
  CHECK(L1.rbloom == L4.rbloom && L1.rt_keys == L4.rt_keys && L1.rt_masks == L4.rt_masks && L1.rt_off == L4.rt_off && L1.rt_size == L4.rt_size,
        "threaded build: rare tables differ");
  CHECK((int64_t)L1.Ubt.size() == L1.n_chunks_pad / 64 * NF2 * 4, "bitmap size");
  CHECK(L1.n_chunks == (n + 31) / 32 && L1.n_chunks_pad % 128 == 0 && L1.n_chunks_pad >= L1.n_chunks, "chunk counts");
  // decode
  std::vector<std::map<uint32_t, uint32_t>> got((size_t)n);
  for (int64_t ch = 0; ch < L1.n_chunks; ch++) {
    const BlockInfo bi = L1.binfo[(size_t)ch];
    const int E = bi.n_entries, E4 = (E + 3) & ~3;
    CHECK(((size_t)bi.off4 * 4 + 2 * (size_t)E4) <= b1.size(), "block %lld outside the array", (long long)ch);
    const uint32_t *words = b1.data() + (size_t)bi.off4 * 4, *masks = words + E4;
    const int rows = (int)std::min<int64_t>(32, n - ch * 32);
    const uint32_t valid = rows == 32 ? 0xFFFFFFFFu : ((1u << rows) - 1u);
    unsigned long long prev = 0;
    std::map<int, uint32_t> umax;
    std::set<int> have2, have2b;
    for (int e = 0; e < E4; e++) {
      if (e >= E) { CHECK(words[e] == PAD_WORD && masks[e] == 0, "padding entry"); continue; }
      const uint32_t w = words[e], f = (w >> 5) & FID_MASK;
      uint32_t t = w & 31u;
      if (t == TF_OVF) t = ovf_get(L1, ch, (uint32_t)e);
      CHECK(t >= 1, "tf of entry");
      const bool rare = e < bi.n_rare, sec = !rare && e < bi.n_rare + bi.n_f2;
      CHECK((L1.fslot[f] < 0 && L1.fslot2[f] == 0xFFFF) == rare && (L1.fslot2[f] != 0xFFFF) == sec,
            "entry %d of chunk %lld in the wrong part of the block", e, (long long)ch);
      const unsigned long long key = ((unsigned long long)f << 16) | t;
      if (e != 0 && e != bi.n_rare && e != bi.n_rare + bi.n_f2) CHECK(key > prev, "entries not sorted");
      prev = key;
      CHECK(masks[e] != 0 && (masks[e] & ~valid) == 0, "mask of entry");
      CHECK(((w & W_ALL) != 0) == (masks[e] == valid), "W_ALL flag");
      if (!rare && !sec) umax[L1.fslot[f]] = std::max(umax[L1.fslot[f]], t);
      CHECK(!(L1.fslot[f] >= 0 && L1.fslot2[f] != 0xFFFF), "feature in both classes");
      if (L1.fslot2[f] != 0xFFFF) { have2.insert((int)L1.fslot2[f]); if (t >= 2) have2b.insert((int)L1.fslot2[f]); }
      if (rare) {  // the block-level inverted index must find the feature, with this chunk in its mask and tf <= its tf
        const int64_t b = ch >> 6;
        const uint32_t bb = rb_bit(f);
        CHECK((L1.rbloom[(size_t)b * (RB_BITS / 32) + (bb >> 5)] >> (bb & 31u)) & 1u, "rare feature missing in the block bitmap");
        const uint32_t size = L1.rt_size[(size_t)b], off = L1.rt_off[(size_t)b];
        uint32_t h = rt_slot(f, size);
        bool found = false;
        for (uint32_t n_probe = 0; n_probe < size; n_probe++) {
          const uint32_t key = L1.rt_keys[off + h];
          if (key == KEY_EMPTY) break;
          if ((key >> 5) == f) {
            found = ((L1.rt_masks[off + h] >> (ch & 63)) & 1ULL) && ((key & 31u) >= std::min<uint32_t>(t, 31u));
            break;
          }
          h = h + 1 == size ? 0 : h + 1;
        }
        CHECK(found, "rare feature %u of chunk %lld not found in its block's table", f, (long long)ch);
      }
      for (int r = 0; r < rows; r++)
        if ((masks[e] >> r) & 1u) {
          auto &row = got[(size_t)c.perm[(size_t)(ch * 32 + r)]];
          CHECK(row.find(f) == row.end(), "feature twice in a row");
          row[f] = t;
        }
    }
    for (int s2 = 0; s2 < NF2; s2++) {
      const bool bit = (L1.Ubt[((size_t)(ch >> 6) * NF2 + s2) * 4 + ((ch & 63) >> 5)] >> (ch & 31)) & 1u;
      const bool bit2 = (L1.Ubt[((size_t)(ch >> 6) * NF2 + s2) * 4 + 2 + ((ch & 63) >> 5)] >> (ch & 31)) & 1u;
      CHECK(bit == (have2.count(s2) != 0) && bit2 == (have2b.count(s2) != 0), "second-class bits of chunk %lld, row %d", (long long)ch, s2);
    }
    for (int s = 0; s < NF; s++) {
      const float u = __half2float(L1.Uf[(size_t)ch * NF + s]);
      auto it = umax.find(s);
      CHECK(u == (it == umax.end() ? 0.f : (float)it->second), "Uf[%lld][%d] = %g", (long long)ch, s, u);
    }
  }
  for (int64_t r = 0; r < n; r++) {
    std::map<uint32_t, uint32_t> want;
    for (int64_t p = c.indptr[(size_t)r]; p < c.indptr[(size_t)r + 1]; p++)
      if (!c.univ[c.ids[(size_t)p]]) want[c.ids[(size_t)p]] = c.tf[(size_t)p];
    CHECK(want == got[(size_t)r], "row %lld does not decode to its features", (long long)r);
  }
  // fixed-point sums: a random query against every row, integer arithmetic as in the kernels vs float64
  std::mt19937 rng(seed + 7);
  std::map<uint32_t, std::pair<double, double>> q;  // feature -> (tf_q a, -d)
  for (int j = 0; j < 40; j++) {
    uint32_t f = (uint32_t)(n_univ + rng() % (V - n_univ));
    q[f] = std::make_pair((1 + rng() % 3) * (1.0 + (rng() % 100000) / 400.0), (rng() % 1000) / 77.0);
  }
  double worst = 0;
  for (int64_t r = 0; r < n; r++) {
    unsigned long long iw = 0, ic = 0;
    double dw = 0, dc = 0;
    for (auto &kv : got[(size_t)r]) {
      auto it = q.find(kv.first);
      if (it == q.end()) continue;
      if (kv.second > 30) continue;  // the 64-bit guard sends queries meeting huge tf to the float64 path
      const unsigned long long w = (unsigned long long)std::llrint(it->second.first * 4294967296.0);
      const unsigned long long cq = (unsigned long long)std::llrint(it->second.second * 16777216.0);
      iw += w * kv.second;
      ic += cq * kv.second * kv.second;
      dw += it->second.first * kv.second;
      dc += it->second.second * kv.second * kv.second;
    }
    if (dw > 0) worst = std::max(worst, std::fabs((double)iw / 4294967296.0 - dw) / dw);
    if (dc > 0) worst = std::max(worst, std::fabs((double)ic / 16777216.0 - dc) / std::max(dc, 1.0));
  }
  CHECK(worst < 1e-7, "fixed-point sums off by %g", worst);
}

int main() {
  check_case(1000, 5000, 3, 1, false);
  check_case(997, 800, 0, 2, true);
  check_case(31, 200, 2, 3, false);
  check_case(4200, 60000, 4, 4, true);
  check_case(1, 50, 0, 5, false);
  // stable_sort_indices == std::stable_sort
  {
    std::mt19937 rng(9);
    std::vector<int> key(50000);
    for (auto &k : key) k = (int)(rng() % 97);
    std::vector<int> a(key.size()), b(key.size());
    for (size_t i = 0; i < a.size(); i++) a[i] = b[i] = (int)i;
    auto less = [&](int x, int y) { return key[(size_t)x] < key[(size_t)y]; };
    stable_sort_indices(a, less, 8);
    std::stable_sort(b.begin(), b.end(), less);
    CHECK(a == b, "parallel index sort differs from std::stable_sort");
  }
  // merge_sorted_runs over uneven, separately sorted runs (what a row-sharded GFKB's ranks exchange) == the full sort
  for (int runs : {1, 2, 3, 5, 8}) {
    std::mt19937 rng(100 + runs);
    std::vector<int> key(70001);
    for (auto &k : key) k = (int)(rng() % 1013);
    auto less = [&](int x, int y) { return key[(size_t)x] < key[(size_t)y]; };
    std::vector<int64_t> cut((size_t)runs + 1, 0);
    for (int r = 1; r < runs; r++) cut[(size_t)r] = (int64_t)(rng() % key.size());
    cut[(size_t)runs] = (int64_t)key.size();
    std::sort(cut.begin(), cut.end());
    std::vector<int> a(key.size()), b(key.size());
    for (size_t i = 0; i < a.size(); i++) a[i] = b[i] = (int)i;
    for (int r = 0; r < runs; r++) std::stable_sort(a.begin() + cut[(size_t)r], a.begin() + cut[(size_t)r + 1], less);
    merge_sorted_runs(a, cut, less, 6);
    std::stable_sort(b.begin(), b.end(), less);
    CHECK(a == b, "merge of sorted runs differs from std::stable_sort");
  }
  if (fails) { printf("%d check(s) failed\n", fails); return 1; }
  printf("all block-builder cases passed\n");
  return 0;
}
