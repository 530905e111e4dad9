// CPU check (no GPU needed; built with nvcc because the headers are CUDA headers) of the host-side batch preparation
// in kakveda_b200/csrc/tile_builder.cuh:
//   * build_tiles_serial must reproduce, byte for byte, the tile builder that ran inside prepare_batch when the GPU
//     parity tests and benchmarks of round 1 were recorded (kept verbatim below as build_tiles_original, taken from
//     git revision 8b6e3da, kakveda_b200/csrc/tfidf_index.cu::prepare_batch);
//   * build_tiles_parallel must equal build_tiles_serial whenever it reports success, and must report failure when a
//     table cap would have closed a tile early;
//   * stable_sort_indices must equal std::stable_sort.
// Exit code 0 = all good.  Run by tests/test_host.py::test_tile_builder_matches_recorded_implementation.
#include "../../kakveda_b200/csrc/tile_builder.cuh"

#include <cstdio>
#include <cstdlib>
#include <random>

int kv_fail(int code, const char *, ...) { return code; }
void kv_clear_error() {}

using namespace kvh;

struct FakeIx {  // the fields the original code read through `ix->`
  int64_t n_total;
  std::vector<uint32_t> h_df;
  int jaccard, corpus_fit;
};

static void build_tiles_original(const FakeIx *ix, const std::vector<QueryPrep> &qp, const std::vector<int> &order,
                                 const std::vector<char> &skip, std::vector<TileDesc> &tiles, std::vector<unsigned char> &tables) {
  const int64_t n_q = (int64_t)order.size();
  const int QT = Tile::QT, H = Tile::H;
  tiles.clear();
  tables.clear();
  // tiles: consecutive sorted queries, closed when 128 queries are in or the feature table is full
  auto new_table = [&]() {
    size_t o = tables.size();
    tables.resize(o + Tile::table_bytes, 0);
    memset(tables.data() + o + Tile::off_keys, 0xFF, sizeof(uint32_t) * H);
  };
  {
    TileDesc cur{0, 0, 0, 0};
    int cur_feats = 0;
    struct Exc0 { uint32_t h, tfq, qi; };
    std::vector<Exc0> exc;                                  // (slot, tf_q > 1, query) of the tile being built
    std::vector<std::pair<uint32_t, uint32_t>> pairs;      // its distinct (feature, tf_q > 1) pairs = extra entries needed
    new_table();
    // extra entries of a finished tile: per feature with exceptions one entry per distinct tf_q value t > 1, holding
    // the weight (t - 1) a(t) and the mask of the queries with exactly that tf_q; the entries of one feature are
    // consecutive (chain flag in .y), the primary slot's key carries KEY_MULTI and the index of the first one
    auto finish_tile = [&](TileDesc &td) {
      unsigned char *tb = tables.data() + tables.size() - Tile::table_bytes;
      uint32_t *keys = (uint32_t *)(tb + Tile::off_keys);
      float *xad = (float *)(tb + Tile::off_xad);
      uint32_t *xmask = (uint32_t *)(tb + Tile::off_xmask);
      std::sort(exc.begin(), exc.end(), [](const Exc0 &a, const Exc0 &b) { return a.h != b.h ? a.h < b.h : (a.tfq != b.tfq ? a.tfq < b.tfq : a.qi < b.qi); });
      int nx = 0;
      for (size_t i = 0; i < exc.size();) {
        const uint32_t h = exc[i].h;
        keys[h] |= KEY_MULTI | ((uint32_t)nx << FID_BITS);
        double a, d;
        idf_host(ix->n_total, ix->h_df[keys[h] & FID_MASK], a, d, ix->jaccard, ix->corpus_fit);
        while (i < exc.size() && exc[i].h == h) {
          const uint32_t t = exc[i].tfq;
          xad[2 * nx] = (float)((double)(t - 1) * a);
          xad[2 * nx + 1] = 1.f;  // another entry of this feature follows (patched below for the last one)
          for (; i < exc.size() && exc[i].h == h && exc[i].tfq == t; i++) xmask[(size_t)nx * TG + (exc[i].qi >> 5)] |= 1u << (exc[i].qi & 31);
          nx++;
        }
        xad[2 * (nx - 1) + 1] = 0.f;
      }
      td.n_extras = nx;
      exc.clear();
      pairs.clear();
    };
    for (int64_t i = 0; i < n_q; i++) {
      const QueryPrep &p = qp[(size_t)order[(size_t)i]];
      unsigned char *tb = tables.data() + tables.size() - Tile::table_bytes;
      uint32_t *keys = (uint32_t *)(tb + Tile::off_keys);
      int fresh = 0, newp = 0;
      if (!skip[(size_t)i]) {
        for (size_t j = 0; j < p.fid.size(); j++) {
          uint32_t f = p.fid[j];
          uint32_t h = (f * 0x9E3779B1u) >> (32 - TLOGH);
          while (keys[h] != KEY_EMPTY && (keys[h] & FID_MASK) != f) h = (h + 1) & (H - 1);
          fresh += keys[h] == KEY_EMPTY;
          if (p.tfq[j] > 1) newp += std::find(pairs.begin(), pairs.end(), std::make_pair(f, p.tfq[j])) == pairs.end();
        }
      }
      if (cur.q_count == QT || cur_feats + fresh > TILE_MAX_FEATURES || (int)pairs.size() + newp > TXCAP) {
        finish_tile(cur);
        tiles.push_back(cur);
        cur = TileDesc{(int)i, 0, 0, 0};
        cur_feats = 0;
        new_table();
        tb = tables.data() + tables.size() - Tile::table_bytes;
        keys = (uint32_t *)(tb + Tile::off_keys);
      }
      const int qi = cur.q_count++;
      if (skip[(size_t)i]) continue;
      float *ad = (float *)(tb + Tile::off_ad);
      uint32_t *masks = (uint32_t *)(tb + Tile::off_masks);
      for (size_t j = 0; j < p.fid.size(); j++) {
        uint32_t f = p.fid[j];
        uint32_t h = (f * 0x9E3779B1u) >> (32 - TLOGH);
        while (keys[h] != KEY_EMPTY && (keys[h] & FID_MASK) != f) h = (h + 1) & (H - 1);
        if (keys[h] == KEY_EMPTY) {
          keys[h] = f;
          double a, d;
          idf_host(ix->n_total, ix->h_df[f], a, d, ix->jaccard, ix->corpus_fit);
          ad[2 * h] = (float)a;
          ad[2 * h + 1] = (float)d;
          cur_feats++;
        }
        masks[(size_t)h * TG + (qi >> 5)] |= 1u << (qi & 31);
        if (p.tfq[j] > 1) {
          exc.push_back(Exc0{h, p.tfq[j], (uint32_t)qi});
          if (std::find(pairs.begin(), pairs.end(), std::make_pair(f, p.tfq[j])) == pairs.end()) pairs.emplace_back(f, p.tfq[j]);
        }
      }
    }
    finish_tile(cur);
    tiles.push_back(cur);
  }
}

static uint64_t rng_state = 0x9E3779B97F4A7C15ULL;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 11); }

// a batch shaped like GFKB queries: a few hundred common features, a Zipf tail of rare ones, some tf > 1, some empty
static void make_batch(int n_q, int vocab, int common, int per_q, double p_multi, double p_skip, std::vector<QueryPrep> &qp,
                       std::vector<char> &skip_by_query) {
  qp.assign((size_t)n_q, QueryPrep());
  skip_by_query.assign((size_t)n_q, 0);
  for (int q = 0; q < n_q; q++) {
    QueryPrep &p = qp[(size_t)q];
    int n = per_q / 2 + (int)(rnd() % (uint32_t)per_q);
    for (int j = 0; j < n; j++) {
      uint32_t f = (rnd() % 10) ? rnd() % (uint32_t)common : rnd() % (uint32_t)vocab;
      if (std::find(p.fid.begin(), p.fid.end(), f) != p.fid.end()) continue;
      p.fid.push_back(f);
      // p_multi < 0.5: only a handful of features repeat inside a query (like "intent" in signature_text), so a tile
      // needs few extra entries; p_multi >= 0.5: any feature may repeat (overflows the 32 extra entries of a tile)
      const bool may = p_multi >= 0.5 || f < 10;
      p.tfq.push_back(may && (rnd() % 1000) < p_multi * 1000 ? 2 + rnd() % 2 : 1);
    }
    p.nq = 1.0 + (rnd() % 100) * 0.1;
    if ((rnd() % 1000) < p_skip * 1000) skip_by_query[(size_t)q] = 1;
  }
}

static int check_case(const char *name, int n_q, int vocab, int common, int per_q, double p_multi, double p_skip, int jaccard,
                      int corpus_fit, bool expect_parallel_ok) {
  FakeIx ix;
  ix.n_total = 1000000;
  ix.jaccard = jaccard;
  ix.corpus_fit = corpus_fit;
  ix.h_df.resize((size_t)vocab);
  for (auto &d : ix.h_df) d = 1 + rnd() % 100000;
  std::vector<QueryPrep> qp;
  std::vector<char> skip_q;
  make_batch(n_q, vocab, common, per_q, p_multi, p_skip, qp, skip_q);
  // order: a stable sort by a coarse key with many ties, through both implementations
  std::vector<int> order((size_t)n_q), order2((size_t)n_q);
  for (int i = 0; i < n_q; i++) order[(size_t)i] = order2[(size_t)i] = i;
  auto less = [&](int a, int b) {
    const int ka = (int)qp[(size_t)a].fid.size() / 4, kb = (int)qp[(size_t)b].fid.size() / 4;
    return ka < kb;
  };
  std::stable_sort(order.begin(), order.end(), less);
  stable_sort_indices(order2, less, 8);
  if (order != order2) { printf("FAIL %s: stable_sort_indices differs from std::stable_sort\n", name); return 1; }
  std::vector<char> skip((size_t)n_q);
  for (int i = 0; i < n_q; i++) skip[(size_t)i] = skip_q[(size_t)order[(size_t)i]];
  const TileCtx cx{ix.n_total, ix.h_df.data(), ix.jaccard, ix.corpus_fit};
  std::vector<TileDesc> t0, t1, t2;
  std::vector<unsigned char> b0, b1, b2;
  build_tiles_original(&ix, qp, order, skip, t0, b0);
  build_tiles_serial(cx, qp, order, skip, t1, b1);
  if (t0.size() != t1.size() || b0 != b1 || memcmp(t0.data(), t1.data(), t0.size() * sizeof(TileDesc)) != 0) {
    printf("FAIL %s: build_tiles_serial differs from the recorded implementation (%zu vs %zu tiles)\n", name, t0.size(), t1.size());
    return 1;
  }
  for (int T : {1, 3, 8}) {
    const bool ok = build_tiles_parallel(cx, qp, order, skip, T, t2, b2);
    if (ok != expect_parallel_ok) { printf("FAIL %s: build_tiles_parallel returned %d with %d threads\n", name, (int)ok, T); return 1; }
    if (ok && (t2.size() != t1.size() || b2 != b1 || memcmp(t2.data(), t1.data(), t1.size() * sizeof(TileDesc)) != 0)) {
      printf("FAIL %s: build_tiles_parallel (%d threads) differs from build_tiles_serial\n", name, T);
      return 1;
    }
  }
  size_t extras = 0;
  for (auto &t : t1) extras += (size_t)t.n_extras;
  printf("ok   %-28s %6d queries %5zu tiles %6zu extra entries parallel=%d\n", name, n_q, t1.size(), extras, (int)expect_parallel_ok);
  return 0;
}

int main() {
  int bad = 0;
  bad += check_case("gfkb-shaped", 20000, 300000, 400, 40, 0.3, 0.01, 0, 0, true);
  bad += check_case("one query", 1, 1000, 50, 10, 0.1, 0.0, 0, 0, true);
  bad += check_case("exactly one tile", 128, 5000, 100, 30, 0.3, 0.0, 0, 0, true);
  bad += check_case("one past a tile", 129, 5000, 100, 30, 0.3, 0.5, 0, 0, true);
  bad += check_case("all skipped", 300, 5000, 100, 30, 0.0, 1.0, 0, 0, true);
  bad += check_case("jaccard weights", 3000, 100000, 300, 50, 0.0, 0.0, 1, 0, true);
  bad += check_case("corpus-fit weights", 3000, 100000, 300, 50, 0.3, 0.0, 0, 1, true);
  bad += check_case("feature cap closes tiles", 4000, 4000000, 2000000, 60, 0.0, 0.0, 0, 0, false);
  bad += check_case("extras cap closes tiles", 4000, 100000, 300, 40, 0.6, 0.0, 0, 0, false);
  if (bad) { printf("%d case(s) failed\n", bad); return 1; }
  printf("all tile-builder cases passed\n");
  return 0;
}
