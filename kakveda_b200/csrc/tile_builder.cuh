// Host-side preparation of a query batch for K1b: text order of the queries and the per-tile shared-memory
// tables (TileLayout in tfidf_kernels.cuh).  Pure host code without CUDA calls, so that it can be unit-tested on a
// box without a GPU (tests/cpp/tile_builder_check.cu, run by tests/test_host.py): the parallel builders below must
// produce byte-for-byte what the sequential reference implementations produce.
#pragma once
#include "tfidf_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>
#include <utility>
#include <vector>

namespace kvh {

using namespace kvk;

// tile shape used by the batched scan: 128 queries, 2048 slots, 32 extra entries (tf_q > 1) per tile
constexpr int TG = 4, TLOGH = 11, TXCAP = 32;
using Tile = TileLayout<TG, TLOGH, TXCAP>;
constexpr int TILE_MAX_FEATURES = (Tile::H * 5) / 8;  // load factor cap 0.625 (linear probing)

template <class F>
void parallel_for(int64_t n, int T, F &&body) {  // body(t, begin, end); thread t owns [n*t/T, n*(t+1)/T)
  T = (int)std::max<int64_t>(1, std::min<int64_t>(T, n));
  std::vector<std::thread> th;
  for (int t = 1; t < T; t++) th.emplace_back([&, t] { body(t, n * t / T, n * (t + 1) / T); });
  body(0, 0, n / T);
  for (auto &x : th) x.join();
}

struct QueryPrep {
  double nq = 0, dotU = 0, corrU = 0, dotS = 0, corrS = 0;  // S: bound start = universal + summary-universal part
  std::vector<uint32_t> fid;  // non-universal, in-vocabulary features
  std::vector<uint32_t> tfq;
};

// idf-derived weights of one feature as every query table uses them (see the header of tfidf_kernels.cuh)
inline void idf_host(int64_t n_total, uint32_t df, double &a, double &d, int jaccard, int corpus_fit) {
  if (jaccard) { a = 1.0; d = 0.0; return; }
  double num = (double)(n_total + (corpus_fit ? 1 : 2));
  double ib = std::log(num / ((double)df + 1.0)) + 1.0;
  double iq = corpus_fit ? ib : std::log(num / ((double)df + 2.0)) + 1.0;
  a = iq * iq;
  d = a - ib * ib;
}

struct TileCtx {  // what the weights depend on
  int64_t n_total;
  const uint32_t *h_df;
  int jaccard, corpus_fit;
};

// ---- index sort: result == std::stable_sort(order = 0..n-1, less) for any strict weak ordering `less` ----
template <class Less>
void stable_sort_indices(std::vector<int> &order, Less less, int T) {
  const int64_t n = (int64_t)order.size();
  auto total = [&](int a, int b) { return less(a, b) || (!less(b, a) && a < b); };  // ties by index == stability
  int parts = 1;
  if (n >= 8192) while (parts * 2 <= T && parts < 64) parts *= 2;
  std::vector<int64_t> cut((size_t)parts + 1);
  for (int i = 0; i <= parts; i++) cut[(size_t)i] = n * i / parts;
  parallel_for(parts, parts, [&](int, int64_t a, int64_t b) {
    for (int64_t i = a; i < b; i++) std::sort(order.begin() + cut[(size_t)i], order.begin() + cut[(size_t)i + 1], total);
  });
  for (int width = 1; width < parts; width *= 2) {
    const int merges = parts / (2 * width);
    parallel_for(merges, merges, [&](int, int64_t a, int64_t b) {
      for (int64_t m = a; m < b; m++)
        std::inplace_merge(order.begin() + cut[(size_t)(m * 2 * width)], order.begin() + cut[(size_t)(m * 2 * width + width)],
                           order.begin() + cut[(size_t)(m * 2 * width + 2 * width)], total);
    });
  }
}

// ---- tile tables ----
struct Exc { uint32_t h, tfq, qi; };

inline void init_table(unsigned char *tb) {
  memset(tb, 0, Tile::table_bytes);
  memset(tb + Tile::off_keys, 0xFF, sizeof(uint32_t) * Tile::H);
}

// extra entries of a finished tile: per feature with exceptions one entry per distinct tf_q value t > 1, holding the
// weight (t - 1) a(t) and the mask of the queries with exactly that tf_q; the entries of one feature are consecutive
// (chain flag in .y), the primary slot's key carries KEY_MULTI and the index of the first one
inline int finish_tile(const TileCtx &cx, unsigned char *tb, std::vector<Exc> &exc) {
  uint32_t *keys = (uint32_t *)(tb + Tile::off_keys);
  float *xad = (float *)(tb + Tile::off_xad);
  uint32_t *xmask = (uint32_t *)(tb + Tile::off_xmask);
  std::sort(exc.begin(), exc.end(), [](const Exc &a, const Exc &b) { return a.h != b.h ? a.h < b.h : (a.tfq != b.tfq ? a.tfq < b.tfq : a.qi < b.qi); });
  int nx = 0;
  for (size_t i = 0; i < exc.size();) {
    const uint32_t h = exc[i].h;
    keys[h] |= KEY_MULTI | ((uint32_t)nx << FID_BITS);
    double a, d;
    idf_host(cx.n_total, cx.h_df[keys[h] & FID_MASK], a, d, cx.jaccard, cx.corpus_fit);
    while (i < exc.size() && exc[i].h == h) {
      const uint32_t t = exc[i].tfq;
      xad[2 * nx] = (float)((double)(t - 1) * a);
      xad[2 * nx + 1] = 1.f;  // another entry of this feature follows (patched below for the last one)
      for (; i < exc.size() && exc[i].h == h && exc[i].tfq == t; i++) xmask[(size_t)nx * TG + (exc[i].qi >> 5)] |= 1u << (exc[i].qi & 31);
      nx++;
    }
    xad[2 * (nx - 1) + 1] = 0.f;
  }
  exc.clear();
  return nx;
}

// add query (sorted slot qi of its tile) to the table; returns the number of features that were new to the table
inline int add_query(const TileCtx &cx, unsigned char *tb, const QueryPrep &p, int qi, std::vector<Exc> &exc,
                     std::vector<std::pair<uint32_t, uint32_t>> &pairs) {
  constexpr int H = Tile::H;
  uint32_t *keys = (uint32_t *)(tb + Tile::off_keys);
  float *ad = (float *)(tb + Tile::off_ad);
  uint32_t *masks = (uint32_t *)(tb + Tile::off_masks);
  int fresh = 0;
  for (size_t j = 0; j < p.fid.size(); j++) {
    uint32_t f = p.fid[j];
    uint32_t h = (f * 0x9E3779B1u) >> (32 - TLOGH);
    while (keys[h] != KEY_EMPTY && (keys[h] & FID_MASK) != f) h = (h + 1) & (H - 1);
    if (keys[h] == KEY_EMPTY) {
      keys[h] = f;
      double a, d;
      idf_host(cx.n_total, cx.h_df[f], a, d, cx.jaccard, cx.corpus_fit);
      ad[2 * h] = (float)a;
      ad[2 * h + 1] = (float)d;
      fresh++;
    }
    masks[(size_t)h * TG + (qi >> 5)] |= 1u << (qi & 31);
    if (p.tfq[j] > 1) {
      exc.push_back(Exc{h, p.tfq[j], (uint32_t)qi});
      if (std::find(pairs.begin(), pairs.end(), std::make_pair(f, p.tfq[j])) == pairs.end()) pairs.emplace_back(f, p.tfq[j]);
    }
  }
  return fresh;
}

// Sequential reference: consecutive sorted queries, a tile is closed when 128 queries are in, or the next query would
// overfill the feature table or the extra entries.  `skip[i]` (by sorted slot): the query takes a slot but no features.
inline void build_tiles_serial(const TileCtx &cx, const std::vector<QueryPrep> &qp, const std::vector<int> &order,
                               const std::vector<char> &skip, std::vector<TileDesc> &tiles, std::vector<unsigned char> &tables) {
  constexpr int QT = Tile::QT, H = Tile::H;
  const int64_t n_q = (int64_t)order.size();
  tiles.clear();
  tables.clear();
  auto new_table = [&]() {
    size_t o = tables.size();
    tables.resize(o + Tile::table_bytes);
    init_table(tables.data() + o);
  };
  TileDesc cur{0, 0, 0, 0};
  int cur_feats = 0;
  std::vector<Exc> exc;
  std::vector<std::pair<uint32_t, uint32_t>> pairs;
  new_table();
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
      cur.n_extras = finish_tile(cx, tb, exc);
      pairs.clear();
      tiles.push_back(cur);
      cur = TileDesc{(int)i, 0, 0, 0};
      cur_feats = 0;
      new_table();
      tb = tables.data() + tables.size() - Tile::table_bytes;
    }
    const int qi = cur.q_count++;
    if (skip[(size_t)i]) continue;
    cur_feats += add_query(cx, tb, p, qi, exc, pairs);
  }
  cur.n_extras = finish_tile(cx, tables.data() + tables.size() - Tile::table_bytes, exc);
  tiles.push_back(cur);
}

// Parallel builder.  Speculates that every tile closes because it holds 128 queries (the table caps are far from
// reached for GFKB-shaped text), builds the tiles independently on T threads and checks the caps as it goes; returns
// false -- leaving the outputs unspecified -- when some tile would have been closed early by the sequential rule, in
// which case the caller runs build_tiles_serial.  When it returns true the outputs are byte-identical to the
// sequential ones (same queries per tile, same insertion order into each table).
inline bool build_tiles_parallel(const TileCtx &cx, const std::vector<QueryPrep> &qp, const std::vector<int> &order,
                                 const std::vector<char> &skip, int T, std::vector<TileDesc> &tiles,
                                 std::vector<unsigned char> &tables) {
  constexpr int QT = Tile::QT;
  const int64_t n_q = (int64_t)order.size();
  const int64_t n_tiles = std::max<int64_t>(1, (n_q + QT - 1) / QT);
  tiles.assign((size_t)n_tiles, TileDesc{0, 0, 0, 0});
  tables.resize((size_t)n_tiles * Tile::table_bytes);
  std::vector<char> bad((size_t)std::max(1, T), 0);
  parallel_for(n_tiles, T, [&](int t, int64_t t0, int64_t t1) {
    std::vector<Exc> exc;
    std::vector<std::pair<uint32_t, uint32_t>> pairs;
    for (int64_t ti = t0; ti < t1 && !bad[(size_t)t]; ti++) {
      unsigned char *tb = tables.data() + (size_t)ti * Tile::table_bytes;
      init_table(tb);
      const int64_t q0 = ti * QT, q1 = std::min<int64_t>(n_q, q0 + QT);
      TileDesc cur{(int)q0, 0, 0, 0};
      int cur_feats = 0;
      exc.clear();
      pairs.clear();
      for (int64_t i = q0; i < q1; i++) {
        const int qi = cur.q_count++;
        if (skip[(size_t)i]) continue;
        cur_feats += add_query(cx, tb, qp[(size_t)order[(size_t)i]], qi, exc, pairs);
        if (cur_feats > TILE_MAX_FEATURES || (int)pairs.size() > TXCAP) { bad[(size_t)t] = 1; break; }
      }
      cur.n_extras = finish_tile(cx, tb, exc);
      tiles[(size_t)ti] = cur;
    }
  });
  for (char b : bad)
    if (b) return false;
  return true;
}

}  // namespace kvh
