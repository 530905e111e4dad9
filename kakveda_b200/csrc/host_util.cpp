// Host-only helpers behind the C ABI that need no device: the global text order of a featurised corpus and the
// gathering of a row subset -- what a rank needs to shard a GFKB by TEXT RANGE instead of by row index
// (kakveda_b200/dist.py, order="text"): similar rows then land in the same shard, so every shard's 64-row chunks are
// as tight as the single index's (row-index sharding spreads near-duplicates over the shards and loosens the block-max
// bounds: 2-GPU pruning 88.6 % vs 91.1 %).
#include "kv_internal.h"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

// lexicographic order of two id sequences (shorter prefix first) -- the second sort key of the scan layout
inline int cmp_ids(const uint32_t *a, int64_t na, const uint32_t *b, int64_t nb) {
  const int64_t n = std::min(na, nb);
  for (int64_t i = 0; i < n; i++)
    if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
  return na < nb ? -1 : (na > nb ? 1 : 0);
}

template <class F>
void run_threads(int64_t n, int T, F &&body) {  // body(begin, end) over [0, n) in T contiguous parts
  T = (int)std::max<int64_t>(1, std::min<int64_t>(T, n));
  std::vector<std::thread> th;
  for (int t = 1; t < T; t++) th.emplace_back([&, t] { body(n * t / T, n * (t + 1) / T); });
  body(0, n / T);
  for (auto &x : th) x.join();
}

}  // namespace

extern "C" {

int kv_text_order(const int64_t *indptr, const uint32_t *ids, int64_t n_rows, int32_t *perm_out, int n_threads) {
  if (n_rows < 0 || n_rows >= (1LL << 31) || (n_rows > 0 && (!indptr || !perm_out)))
    return kv_fail(KV_ERR_INVALID, "kv_text_order: bad arguments");
  for (int64_t i = 0; i < n_rows; i++)
    if (indptr[i + 1] < indptr[i]) return kv_fail(KV_ERR_INVALID, "kv_text_order: indptr not monotone");
  if (n_rows > 0 && indptr[n_rows] > indptr[0] && !ids) return kv_fail(KV_ERR_INVALID, "kv_text_order: ids is NULL");
  int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  T = std::max(1, std::min(T, 64));
  auto less = [&](int32_t a, int32_t b) {
    const int c = cmp_ids(ids + indptr[a], indptr[a + 1] - indptr[a], ids + indptr[b], indptr[b + 1] - indptr[b]);
    return c != 0 ? c < 0 : a < b;  // equal text: lower row first
  };
  for (int64_t i = 0; i < n_rows; i++) perm_out[i] = (int32_t)i;
  int parts = 1;
  if (n_rows >= 50000) while (parts * 2 <= T) parts *= 2;
  std::vector<int64_t> cut((size_t)parts + 1);
  for (int i = 0; i <= parts; i++) cut[(size_t)i] = n_rows * i / parts;
  run_threads(parts, parts, [&](int64_t a, int64_t b) {
    for (int64_t i = a; i < b; i++) std::sort(perm_out + cut[(size_t)i], perm_out + cut[(size_t)i + 1], less);
  });
  for (int width = 1; width < parts; width *= 2) {
    const int merges = parts / (2 * width);
    run_threads(merges, merges, [&](int64_t a, int64_t b) {
      for (int64_t m = a; m < b; m++)
        std::inplace_merge(perm_out + cut[(size_t)(m * 2 * width)], perm_out + cut[(size_t)(m * 2 * width + width)],
                           perm_out + cut[(size_t)(m * 2 * width + 2 * width)], less);
    });
  }
  return KV_OK;
}

int kv_csr_gather_rows(const int64_t *indptr, const uint32_t *ids, const uint32_t *tf, int64_t n_rows, const int64_t *rows,
                       int64_t n_sel, const int64_t *out_indptr, uint32_t *out_ids, uint32_t *out_tf, int n_threads) {
  if (n_sel < 0 || n_rows < 0 || (n_sel > 0 && (!indptr || !rows || !out_indptr)))
    return kv_fail(KV_ERR_INVALID, "kv_csr_gather_rows: bad arguments");
  for (int64_t i = 0; i < n_sel; i++) {
    if (rows[i] < 0 || rows[i] >= n_rows) return kv_fail(KV_ERR_INVALID, "kv_csr_gather_rows: row %lld outside 0..%lld", (long long)rows[i], (long long)n_rows);
    const int64_t len = indptr[rows[i] + 1] - indptr[rows[i]];
    if (len < 0 || out_indptr[i + 1] - out_indptr[i] != len)
      return kv_fail(KV_ERR_INVALID, "kv_csr_gather_rows: out_indptr does not match the selected rows (row %lld)", (long long)rows[i]);
    if (len > 0 && (!ids || !tf || !out_ids || !out_tf)) return kv_fail(KV_ERR_INVALID, "kv_csr_gather_rows: NULL array");
  }
  int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  T = std::max(1, std::min(T, 64));
  run_threads(n_sel, n_sel >= 4096 ? T : 1, [&](int64_t a, int64_t b) {
    for (int64_t i = a; i < b; i++) {
      const int64_t src = indptr[rows[i]], len = indptr[rows[i] + 1] - src, dst = out_indptr[i];
      if (len > 0) {
        memcpy(out_ids + dst, ids + src, (size_t)len * sizeof(uint32_t));
        memcpy(out_tf + dst, tf + src, (size_t)len * sizeof(uint32_t));
      }
    }
  });
  return KV_OK;
}

}  // extern "C"
