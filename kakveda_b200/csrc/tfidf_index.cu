// TF-IDF cosine index for the GFKB match path: the kv_index handle and its C ABI
// (include/kakveda_b200.h).  Device code lives in tfidf_kernels.cuh; see there for the math and
// the HBM layout.  Host responsibilities: keep the append-only CSR, finalize (statistics on the
// device, text-order sort of the rows and chunk summaries on the host cores), per-batch query
// preparation (per-query constants in float64, tile tables) and kernel launches.
#include "tfidf_kernels.cuh"
#include "tile_builder.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iterator>
#include <limits>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

using namespace kvk;
using namespace kvh;  // tile shape, QueryPrep, idf_host, parallel_for, tile builders (tile_builder.cuh)

namespace {
int host_threads() {
  int t = (int)std::thread::hardware_concurrency();
  if (const char *e = getenv("KAKVEDA_B200_THREADS")) t = atoi(e);
  return std::max(1, std::min(t, 64));
}

}  // namespace

// ----------------------------------------------------------------------------------------
// handle
// ----------------------------------------------------------------------------------------
struct kv_index {
  int device = 0;
  int64_t row_base = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  std::mutex mu;
  int sm_count = 148;

  // raw CSR: device copy for the statistics kernels, host copy for the sort and the summaries
  DevVec<int64_t> indptr;  // n_rows + 1 entries once any row exists
  DevVec<uint32_t> ids;
  DevVec<uint16_t> tf;
  std::vector<int64_t> h_indptr{0};
  std::vector<uint32_t> h_ids;
  std::vector<uint16_t> h_tf;
  int64_t n_rows = 0, nnz = 0;

  bool has_gdf = false;
  std::vector<uint32_t> h_gdf;
  int64_t n_global = 0;

  bool finalized = false;
  int jaccard = 0;  // 0: TF-IDF cosine (the reference's measure), 1: token-set Jaccard (K3)
  int corpus_fit = 0;  // 1: TF-IDF fitted on the corpus only (self-join / pattern clustering); the query is just transformed
  int64_t V = 0, n_total = 0;
  DevBuf<uint32_t> d_df, d_cnt, d_tfmin, d_tfmax, d_utf;
  DevBuf<double> d_a64, d_d64, d_bb64, d_B64;
  DevBuf<float> d_B32, d_cminB;
  DevBuf<uint8_t> d_univ;
  DevBuf<int> d_perm, d_invperm;
  // scan layout currently on the device (perm, stream, summaries): which rows / universal set it was built for
  bool layout_valid = false;
  int64_t layout_rows = -1;
  std::vector<uint8_t> layout_univ;
  int last_finalize_kind = 0;  // 1: full rebuild, 2: statistics-only refresh
  // K6 scratch
  DevBuf<int64_t> d_rq_indptr;
  DevBuf<uint32_t> d_rq_ids, d_rq_tf;
  DevBuf<double> d_rq_const, d_rq_out;
  DevBuf<long long> d_rq_rows;
  DevBuf<int64_t> d_chunkptr, d_sumptr, d_grpptr;
  DevBuf<uint32_t> d_stream, d_sum_stream, d_grp_stream;
  DevBuf<unsigned long long> d_ovf_keys;
  DevBuf<uint32_t> d_ovf_vals;
  int n_ovf = 0;
  int64_t stream_len = 0, sum_len = 0, grp_len = 0, n_chunks = 0;
  std::vector<uint32_t> h_df;
  std::vector<uint8_t> h_univ;
  std::vector<uint32_t> h_utf;
  std::vector<std::pair<uint32_t, uint32_t>> h_su;  // summary-universal features (fid, tf), sorted
  int64_t n_univ = 0;

  // query scratch
  PinnedBuf<unsigned char> h_tables;
  PinnedBuf<TileDesc> h_tiles;
  PinnedBuf<float> h_qconst;  // 4 * n_q
  PinnedBuf<int> h_qperm;     // 2 * n_q: sorted slot -> original query, then null-query list
  DevBuf<unsigned char> d_tables;
  DevBuf<TileDesc> d_tiles;
  DevBuf<float> d_qconst;
  DevBuf<int> d_qperm;
  DevBuf<int> d_gthr;
  // cross-GPU threshold exchange (row-sharded GFKB): d_gthr is exported over CUDA IPC, the peers' arrays are mapped here
  bool gthr_exported = false;
  int n_peers = 0;
  int *peer_gthr[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int64_t peer_cap = 0;  // queries the exchanged arrays hold
  DevBuf<int> d_excl_sorted, d_excl_orig;  // self-join exclusions of the resident batch (by sorted slot / by original query)
  std::vector<int> h_excl_orig;
  bool has_excl = false;
  DevBuf<float> d_ubuf;
  DevBuf<unsigned long long> d_stats;
  DevBuf<float> d_part_s, d_out_s;
  DevBuf<long long> d_part_r, d_out_r;
  PinnedBuf<float> h_out_s;
  PinnedBuf<long long> h_out_r;
  // single-query scratch
  PinnedBuf<unsigned char> h_qtab;
  DevBuf<unsigned char> d_qtab;
  DevBuf<double> d_scores;

  // query batch currently resident on the device (kv_query_upload / first half of kv_topk)
  bool batch_valid = false;
  int64_t batch_q = 0, batch_tiles = 0, batch_h2d_bytes = 0, batch_null = 0;
  std::vector<int64_t> irr_q, irr_indptr;
  std::vector<uint32_t> irr_ids, irr_tf;
  std::vector<double> irr_oov;

  float last_ms[4] = {0, 0, 0, 0};
  float last_score_ms = 0;
  int64_t last_ctas = 0, last_tiles = 0, last_splits = 0;
  unsigned long long last_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

static void close_peers(kv_index *ix) {
  for (int i = 0; i < ix->n_peers; i++)
    if (ix->peer_gthr[i]) cudaIpcCloseMemHandle(ix->peer_gthr[i]);
  for (auto &p : ix->peer_gthr) p = nullptr;
  ix->n_peers = 0;
  ix->peer_cap = 0;
}

namespace {

void prep_query(const kv_index *ix, const uint32_t *ids, const uint32_t *tf, int64_t nnz, double oov_tf2,
                QueryPrep &out) {
  out.dotU = out.corrU = out.dotS = out.corrS = 0;
  out.fid.clear();
  out.tfq.clear();
  // df == 0 features of the query: the refit gives them idf ln((N+2)/2)+1; a corpus-only fit does not know them at all
  double idf0 = ix->jaccard ? 1.0 : (ix->corpus_fit ? 0.0 : std::log((double)(ix->n_total + 2) / 2.0) + 1.0);
  out.nq = oov_tf2 * idf0 * idf0;
  for (int64_t i = 0; i < nnz; i++) {
    uint32_t t = ids[i];
    double f = (double)tf[i];
    if ((int64_t)t >= ix->V) {  // id issued after finalize: in no indexed row
      out.nq += f * f * idf0 * idf0;
      continue;
    }
    double a, d;
    idf_host(ix->n_total, ix->h_df[t], a, d, ix->jaccard, ix->corpus_fit);
    out.nq += f * f * a;
    if (ix->h_univ[t]) {
      double u = (double)ix->h_utf[t];
      out.dotU += f * u * a;
      out.corrU += u * u * d;
    } else {
      out.fid.push_back(t);
      out.tfq.push_back(tf[i]);
      auto it = std::lower_bound(ix->h_su.begin(), ix->h_su.end(), std::make_pair(t, 0u));
      if (it != ix->h_su.end() && it->first == t) {
        double u = (double)it->second;
        out.dotS += f * u * a;
        out.corrS += u * u * d;
      }
    }
  }
  out.dotS += out.dotU;
  out.corrS += out.corrU;
}

// lexicographic order of two id sequences (shorter prefix first)
inline int cmp_seq(const uint32_t *a, int64_t na, const uint32_t *b, int64_t nb) {
  int64_t n = std::min(na, nb);
  for (int64_t i = 0; i < n; i++)
    if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
  return na < nb ? -1 : (na > nb ? 1 : 0);
}

// perm := row indices sorted by (norm class, feature-id sequence).  The featuriser emits 1-grams in
// token order first, so the second key is text order up to the naming of tokens: rows with similar
// text become neighbours.  The first key groups rows whose norms B_c lie within a factor sqrt(2),
// which keeps the chunk bound (it uses the smallest norm of the chunk) tight.  Ties keep the
// original order.
void sort_rows_by_text(const std::vector<int64_t> &indptr, const std::vector<uint32_t> &ids, int64_t n,
                       const std::vector<float> &B, std::vector<int> &perm) {
  perm.resize((size_t)n);
  std::vector<short> cls((size_t)n);
  for (int64_t i = 0; i < n; i++) {
    perm[(size_t)i] = (int)i;
    cls[(size_t)i] = B[(size_t)i] > 0.f ? (short)std::floor(std::log2((double)B[(size_t)i]) * 2.0) : (short)-1000;
  }
  auto less = [&](int a, int b) {
    if (cls[(size_t)a] != cls[(size_t)b]) return cls[(size_t)a] < cls[(size_t)b];
    int c = cmp_seq(ids.data() + indptr[a], indptr[a + 1] - indptr[a], ids.data() + indptr[b],
                    indptr[b + 1] - indptr[b]);
    return c != 0 ? c < 0 : a < b;
  };
  int T = host_threads();
  if (n < 50000) T = 1;
  int parts = 1;
  while (parts * 2 <= T) parts *= 2;
  std::vector<int64_t> cut((size_t)parts + 1);
  for (int i = 0; i <= parts; i++) cut[(size_t)i] = n * i / parts;
  parallel_for(parts, parts, [&](int, int64_t a, int64_t b) {
    for (int64_t i = a; i < b; i++) std::sort(perm.begin() + cut[(size_t)i], perm.begin() + cut[(size_t)i + 1], less);
  });
  for (int width = 1; width < parts; width *= 2) {
    int merges = parts / (2 * width);
    parallel_for(merges, merges, [&](int, int64_t a, int64_t b) {
      for (int64_t m = a; m < b; m++) {
        int64_t lo = cut[(size_t)(m * 2 * width)], mid = cut[(size_t)(m * 2 * width + width)],
                hi = cut[(size_t)(m * 2 * width + 2 * width)];
        std::inplace_merge(perm.begin() + lo, perm.begin() + mid, perm.begin() + hi, less);
      }
    });
  }
}

}  // namespace

extern "C" {

int kv_index_create(int device, int64_t row_base, kv_index **out) {
  if (!out) return kv_fail(KV_ERR_INVALID, "kv_index_create: out is NULL");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return kv_fail(KV_ERR_CUDA, "kv_index_create: no CUDA device visible (this library has no CPU path)");
  }
  if (device < 0 || device >= n) return kv_fail(KV_ERR_INVALID, "kv_index_create: device %d out of range", device);
  KV_CUDA(cudaSetDevice(device));
  kv_index *ix = new kv_index();
  ix->device = device;
  ix->row_base = row_base;
  cudaDeviceProp prop;
  KV_CUDA(cudaGetDeviceProperties(&prop, device));
  ix->sm_count = prop.multiProcessorCount;
  KV_CUDA(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));
  for (auto &e : ix->ev) KV_CUDA(cudaEventCreate(&e));
  *out = ix;
  return KV_OK;
}

void kv_index_destroy(kv_index *ix) {
  if (!ix) return;
  cudaSetDevice(ix->device);
  cudaStreamSynchronize(ix->stream);
  ix->indptr.release(); ix->ids.release(); ix->tf.release();
  ix->d_df.release(); ix->d_cnt.release(); ix->d_tfmin.release(); ix->d_tfmax.release(); ix->d_utf.release();
  ix->d_a64.release(); ix->d_d64.release(); ix->d_bb64.release(); ix->d_B64.release();
  ix->d_B32.release(); ix->d_cminB.release(); ix->d_univ.release(); ix->d_perm.release();
  ix->d_chunkptr.release(); ix->d_sumptr.release(); ix->d_grpptr.release(); ix->d_grp_stream.release();
  ix->d_stream.release(); ix->d_sum_stream.release();
  ix->d_ovf_keys.release(); ix->d_ovf_vals.release();
  ix->h_tables.release(); ix->h_tiles.release(); ix->h_qconst.release(); ix->h_qperm.release();
  ix->d_tables.release(); ix->d_tiles.release(); ix->d_qconst.release(); ix->d_qperm.release(); ix->d_gthr.release();
  close_peers(ix);
  ix->d_excl_sorted.release(); ix->d_excl_orig.release(); ix->d_invperm.release();
  ix->d_rq_indptr.release(); ix->d_rq_ids.release(); ix->d_rq_tf.release(); ix->d_rq_const.release(); ix->d_rq_out.release();
  ix->d_rq_rows.release();
  ix->d_ubuf.release(); ix->d_stats.release();
  ix->d_part_s.release(); ix->d_out_s.release(); ix->d_part_r.release(); ix->d_out_r.release();
  ix->h_out_s.release(); ix->h_out_r.release();
  ix->h_qtab.release(); ix->d_qtab.release(); ix->d_scores.release();
  for (auto &e : ix->ev) if (e) cudaEventDestroy(e);
  if (ix->stream) cudaStreamDestroy(ix->stream);
  delete ix;
}

int64_t kv_index_rows(const kv_index *ix) { return ix ? ix->n_rows : 0; }

int kv_index_append(kv_index *ix, const int64_t *indptr, const uint32_t *ids, const uint32_t *tf, int64_t n_rows) {
  if (!ix || n_rows < 0 || (n_rows > 0 && !indptr)) return kv_fail(KV_ERR_INVALID, "kv_index_append: bad arguments");
  if (n_rows == 0) return KV_OK;
  std::lock_guard<std::mutex> g(ix->mu);
  KV_CUDA(cudaSetDevice(ix->device));
  const int64_t add = indptr[n_rows] - indptr[0];
  if (add < 0 || (add > 0 && (!ids || !tf))) return kv_fail(KV_ERR_INVALID, "kv_index_append: bad CSR");
  if (ix->n_rows + n_rows >= (1LL << 31) - CHUNK_ROWS)
    return kv_fail(KV_ERR_INVALID, "kv_index_append: more than 2^31 rows in one shard");
  for (int64_t i = 1; i <= n_rows; i++)
    if (indptr[i] < indptr[i - 1]) return kv_fail(KV_ERR_INVALID, "kv_index_append: indptr not monotone");
  const uint32_t *tfs = tf + indptr[0];
  for (int64_t i = 0; i < add; i++)
    if (tfs[i] == 0 || tfs[i] > 65535u || (ix->jaccard && tfs[i] != 1))
      return kv_fail(KV_ERR_INVALID, "kv_index_append: term frequency %u outside 1..65535 (or != 1 in Jaccard mode)", tfs[i]);
  try {
    ix->h_indptr.reserve((size_t)(ix->n_rows + n_rows + 1));
    for (int64_t i = 1; i <= n_rows; i++) ix->h_indptr.push_back(indptr[i] - indptr[0] + ix->nnz);
    ix->h_ids.insert(ix->h_ids.end(), ids + indptr[0], ids + indptr[0] + add);
    ix->h_tf.resize((size_t)(ix->nnz + add));
    for (int64_t i = 0; i < add; i++) ix->h_tf[(size_t)(ix->nnz + i)] = (uint16_t)tfs[i];
  } catch (const std::bad_alloc &) {
    return kv_fail(KV_ERR_NOMEM, "kv_index_append: out of host memory");
  }
  KV_CUDA(ix->indptr.reserve(ix->n_rows + n_rows + 1, ix->stream));
  KV_CUDA(ix->ids.reserve(ix->nnz + add, ix->stream));
  KV_CUDA(ix->tf.reserve(ix->nnz + add, ix->stream));
  KV_CUDA(cudaMemcpyAsync(ix->indptr.p + ix->n_rows, ix->h_indptr.data() + ix->n_rows,
                          (size_t)(n_rows + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ix->stream));
  if (add) {
    KV_CUDA(cudaMemcpyAsync(ix->ids.p + ix->nnz, ix->h_ids.data() + ix->nnz, (size_t)add * sizeof(uint32_t),
                            cudaMemcpyHostToDevice, ix->stream));
    KV_CUDA(cudaMemcpyAsync(ix->tf.p + ix->nnz, ix->h_tf.data() + ix->nnz, (size_t)add * sizeof(uint16_t),
                            cudaMemcpyHostToDevice, ix->stream));
  }
  KV_CUDA(cudaStreamSynchronize(ix->stream));
  ix->n_rows += n_rows;
  ix->nnz += add;
  ix->indptr.n = ix->n_rows + 1;
  ix->ids.n = ix->nnz;
  ix->tf.n = ix->nnz;
  ix->finalized = false;
  return KV_OK;
}

int kv_index_set_mode(kv_index *ix, int mode) {
  if (!ix || (mode != KV_MODE_TFIDF_COSINE && mode != KV_MODE_JACCARD && mode != KV_MODE_TFIDF_CORPUS_FIT))
    return kv_fail(KV_ERR_INVALID, "kv_index_set_mode: mode must be KV_MODE_TFIDF_COSINE, KV_MODE_JACCARD or KV_MODE_TFIDF_CORPUS_FIT");
  std::lock_guard<std::mutex> g(ix->mu);
  if (mode == KV_MODE_JACCARD)
    for (uint16_t f : ix->h_tf)
      if (f != 1) return kv_fail(KV_ERR_INVALID, "kv_index_set_mode: Jaccard rows are token SETS (every tf must be 1)");
  ix->jaccard = mode == KV_MODE_JACCARD;
  ix->corpus_fit = mode == KV_MODE_TFIDF_CORPUS_FIT;
  ix->finalized = false;
  ix->layout_valid = false;
  return KV_OK;
}

// |q ∩ row| and |q ∪ row| of already selected (query, row) pairs, exact integers (host; Q*k pairs)
int kv_jaccard_counts(kv_index *ix, const int64_t *q_indptr, const uint32_t *q_ids, const double *q_oov_tf2, int64_t n_q,
                      int k, const int64_t *rows, int32_t *out_inter, int32_t *out_union) {
  if (!ix || n_q < 0 || k < 1 || (n_q > 0 && (!q_indptr || !rows || !out_inter || !out_union)))
    return kv_fail(KV_ERR_INVALID, "kv_jaccard_counts: bad arguments");
  std::lock_guard<std::mutex> g(ix->mu);
  std::vector<uint32_t> qs, rs;
  for (int64_t q = 0; q < n_q; q++) {
    qs.assign(q_ids + q_indptr[q], q_ids + q_indptr[q + 1]);
    std::sort(qs.begin(), qs.end());
    const int64_t nq = (int64_t)qs.size() + (q_oov_tf2 ? (int64_t)q_oov_tf2[q] : 0);
    for (int j = 0; j < k; j++) {
      const int64_t r = rows[q * k + j] - ix->row_base;
      if (rows[q * k + j] < 0 || r < 0 || r >= ix->n_rows) { out_inter[q * k + j] = out_union[q * k + j] = -1; continue; }
      rs.assign(ix->h_ids.begin() + ix->h_indptr[(size_t)r], ix->h_ids.begin() + ix->h_indptr[(size_t)r + 1]);
      std::sort(rs.begin(), rs.end());
      int32_t inter = 0;
      for (size_t a = 0, b = 0; a < qs.size() && b < rs.size();) {
        if (qs[a] == rs[b]) { inter++; a++; b++; }
        else if (qs[a] < rs[b]) a++;
        else b++;
      }
      out_inter[q * k + j] = inter;
      out_union[q * k + j] = (int32_t)(nq + (int64_t)rs.size() - inter);
    }
  }
  return KV_OK;
}

int kv_index_set_global_df(kv_index *ix, const uint32_t *df, int64_t vocab_size, int64_t n_rows_global) {
  if (!ix || !df || vocab_size < 0 || n_rows_global < 0)
    return kv_fail(KV_ERR_INVALID, "kv_index_set_global_df: bad arguments");
  std::lock_guard<std::mutex> g(ix->mu);
  ix->h_gdf.assign(df, df + vocab_size);
  ix->n_global = n_rows_global;
  ix->has_gdf = true;
  ix->finalized = false;
  return KV_OK;
}

int kv_index_local_df(kv_index *ix, uint32_t *df_out, int64_t vocab_size) {
  if (!ix || !df_out || vocab_size < 0) return kv_fail(KV_ERR_INVALID, "kv_index_local_df: bad arguments");
  std::lock_guard<std::mutex> g(ix->mu);
  KV_CUDA(cudaSetDevice(ix->device));
  for (uint32_t t : ix->h_ids)
    if ((int64_t)t >= vocab_size) return kv_fail(KV_ERR_INVALID, "kv_index_local_df: feature id %u outside vocabulary", t);
  KV_CUDA(ix->d_cnt.ensure(std::max<int64_t>(vocab_size, 1)));
  KV_CUDA(cudaMemsetAsync(ix->d_cnt.p, 0, (size_t)std::max<int64_t>(vocab_size, 1) * sizeof(uint32_t), ix->stream));
  if (ix->nnz) {
    hist_kernel<<<ix->sm_count * 8, 256, 0, ix->stream>>>(ix->ids.p, ix->tf.p, ix->nnz, ix->d_cnt.p, nullptr, nullptr);
    KV_CUDA(cudaGetLastError());
  }
  if (vocab_size)
    KV_CUDA(cudaMemcpyAsync(df_out, ix->d_cnt.p, (size_t)vocab_size * sizeof(uint32_t), cudaMemcpyDeviceToHost, ix->stream));
  KV_CUDA(cudaStreamSynchronize(ix->stream));
  return KV_OK;
}

// a(t), d(t) as the HOST computes them (idf_host: the values every query table is built from) -> d_a64/d_d64, so
// that K6 multiplies the very same doubles as K1a
static int upload_idf_tables(kv_index *ix, int64_t V) {
  if (V <= 0) return KV_OK;
  std::vector<double> ha((size_t)V), hd((size_t)V);
  parallel_for(V, V >= 65536 ? host_threads() : 1, [&](int, int64_t a0, int64_t a1) {
    for (int64_t t = a0; t < a1; t++) idf_host(ix->n_total, ix->h_df[(size_t)t], ha[(size_t)t], hd[(size_t)t], ix->jaccard, ix->corpus_fit);
  });
  KV_CUDA(cudaMemcpyAsync(ix->d_a64.p, ha.data(), (size_t)V * 8, cudaMemcpyHostToDevice, ix->stream));
  KV_CUDA(cudaMemcpyAsync(ix->d_d64.p, hd.data(), (size_t)V * 8, cudaMemcpyHostToDevice, ix->stream));
  KV_CUDA(cudaStreamSynchronize(ix->stream));
  return KV_OK;
}

int kv_index_finalize(kv_index *ix, int64_t vocab_size) {
  if (!ix || vocab_size < 0) return kv_fail(KV_ERR_INVALID, "kv_index_finalize: bad arguments");
  std::lock_guard<std::mutex> g(ix->mu);
  KV_CUDA(cudaSetDevice(ix->device));
  if (vocab_size >= (int64_t)FID_NONE)
    return kv_fail(KV_ERR_INVALID, "kv_index_finalize: vocabulary of %lld features exceeds the 2^26-1 the scan layout encodes",
                   (long long)vocab_size);
  if (ix->has_gdf && (int64_t)ix->h_gdf.size() != vocab_size)
    return kv_fail(KV_ERR_INVALID, "kv_index_finalize: global df has %lld entries, vocabulary %lld",
                   (long long)ix->h_gdf.size(), (long long)vocab_size);
  cudaStream_t s = ix->stream;
  const int64_t V = vocab_size, n = ix->n_rows;
  {
    uint32_t mx = 0;
    for (uint32_t t : ix->h_ids) mx = std::max(mx, t);
    if (ix->nnz && (int64_t)mx >= V)
      return kv_fail(KV_ERR_INVALID, "kv_index_finalize: feature id %u outside vocabulary of %lld", mx, (long long)V);
  }
  // ---- statistics on the device: df, idf tables, universal features ----
  const int64_t Vz = V > 0 ? V : 1;
  KV_CUDA(ix->d_df.ensure(Vz)); KV_CUDA(ix->d_cnt.ensure(Vz)); KV_CUDA(ix->d_tfmin.ensure(Vz));
  KV_CUDA(ix->d_tfmax.ensure(Vz)); KV_CUDA(ix->d_utf.ensure(Vz));
  KV_CUDA(ix->d_a64.ensure(Vz)); KV_CUDA(ix->d_d64.ensure(Vz)); KV_CUDA(ix->d_bb64.ensure(Vz));
  KV_CUDA(ix->d_univ.ensure(Vz));
  KV_CUDA(cudaMemsetAsync(ix->d_cnt.p, 0, (size_t)Vz * 4, s));
  KV_CUDA(cudaMemsetAsync(ix->d_tfmin.p, 0xFF, (size_t)Vz * 4, s));
  KV_CUDA(cudaMemsetAsync(ix->d_tfmax.p, 0, (size_t)Vz * 4, s));
  if (ix->nnz) {
    hist_kernel<<<ix->sm_count * 8, 256, 0, s>>>(ix->ids.p, ix->tf.p, ix->nnz, ix->d_cnt.p, ix->d_tfmin.p, ix->d_tfmax.p);
    KV_CUDA(cudaGetLastError());
  }
  if (ix->has_gdf) {
    if (V) KV_CUDA(cudaMemcpyAsync(ix->d_df.p, ix->h_gdf.data(), (size_t)V * 4, cudaMemcpyHostToDevice, s));
    ix->n_total = ix->n_global;
  } else {
    if (V) KV_CUDA(cudaMemcpyAsync(ix->d_df.p, ix->d_cnt.p, (size_t)V * 4, cudaMemcpyDeviceToDevice, s));
    ix->n_total = n;
  }
  if (V) {
    IdfTables T{ix->d_a64.p, ix->d_d64.p, ix->d_bb64.p, ix->d_univ.p, ix->d_utf.p};
    idf_kernel<<<(unsigned)((V + 255) / 256), 256, 0, s>>>(ix->d_df.p, ix->d_cnt.p, ix->d_tfmin.p, ix->d_tfmax.p, V,
                                                            ix->n_total, n, ix->jaccard, ix->corpus_fit, T);
    KV_CUDA(cudaGetLastError());
  }
  ix->h_df.assign((size_t)V, 0);
  ix->h_univ.assign((size_t)V, 0);
  ix->h_utf.assign((size_t)V, 0);
  if (V) {
    KV_CUDA(cudaMemcpyAsync(ix->h_df.data(), ix->d_df.p, (size_t)V * 4, cudaMemcpyDeviceToHost, s));
    KV_CUDA(cudaMemcpyAsync(ix->h_univ.data(), ix->d_univ.p, (size_t)V, cudaMemcpyDeviceToHost, s));
    KV_CUDA(cudaMemcpyAsync(ix->h_utf.data(), ix->d_utf.p, (size_t)V * 4, cudaMemcpyDeviceToHost, s));
  }
  // row norms in original order (the sort key needs them)
  const int64_t nz = n > 0 ? n : 1;
  KV_CUDA(ix->d_perm.ensure(nz));
  KV_CUDA(ix->d_B64.ensure(nz)); KV_CUDA(ix->d_B32.ensure(nz));
  std::vector<float> hB((size_t)n);
  if (n) {
    rownorm_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, s>>>(ix->indptr.p, ix->ids.p, ix->tf.p, nullptr, n,
                                                                    ix->d_bb64.p, ix->d_univ.p, ix->d_B64.p,
                                                                    ix->d_B32.p, nullptr);
    KV_CUDA(cudaGetLastError());
    KV_CUDA(cudaMemcpyAsync(hB.data(), ix->d_B32.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, s));
  }
  KV_CUDA(cudaStreamSynchronize(s));
  ix->n_univ = 0;
  for (uint8_t u : ix->h_univ) ix->n_univ += u;
  {
    int rc = upload_idf_tables(ix, V);
    if (rc != KV_OK) return rc;
  }
  // ---- statistics-only refresh: the rows (hence text order, stream, chunk summaries) are the ones the device layout
  // was built from and the set of folded universal features is unchanged; only N / df moved (rows were appended to
  // ANOTHER shard or segment of the same GFKB).  Row norms and chunk minima are recomputed, nothing is re-sorted.
  if (ix->layout_valid && ix->layout_rows == n && n > 0 && !getenv("KAKVEDA_B200_FULL_FINALIZE")) {
    bool same = (int64_t)ix->layout_univ.size() <= V;
    for (size_t t = 0; same && t < ix->layout_univ.size(); t++) same = ix->layout_univ[t] == ix->h_univ[t];
    for (size_t t = ix->layout_univ.size(); same && t < (size_t)V; t++) same = ix->h_univ[t] == 0;
    if (same) {
      rownorm_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, s>>>(ix->indptr.p, ix->ids.p, ix->tf.p, ix->d_perm.p, n,
                                                                      ix->d_bb64.p, ix->d_univ.p, ix->d_B64.p,
                                                                      ix->d_B32.p, nullptr);
      KV_CUDA(cudaGetLastError());
      chunk_meta_kernel<<<(unsigned)((ix->n_chunks + 255) / 256), 256, 0, s>>>(ix->d_B32.p, n, ix->n_chunks, ix->d_cminB.p);
      KV_CUDA(cudaGetLastError());
      KV_CUDA(cudaStreamSynchronize(s));
      ix->V = V;
      ix->finalized = true;
      ix->batch_valid = false;
      ix->last_finalize_kind = 2;
      return KV_OK;
    }
  }
  ix->layout_valid = false;
  // ---- on the host cores: (norm class, text) order of the rows ----
  std::vector<int> perm;
  sort_rows_by_text(ix->h_indptr, ix->h_ids, n, hB, perm);
  ix->n_chunks = (n + CHUNK_ROWS - 1) / CHUNK_ROWS;
  KV_CUDA(ix->d_chunkptr.ensure(ix->n_chunks + 1));
  KV_CUDA(ix->d_sumptr.ensure(ix->n_chunks + 1));
  KV_CUDA(ix->d_cminB.ensure(ix->n_chunks + 1));
  ix->stream_len = ix->sum_len = 0;
  ix->n_ovf = 0;
  if (n) {
    KV_CUDA(cudaMemcpyAsync(ix->d_perm.p, perm.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, s));
    KV_CUDA(ix->d_invperm.ensure(n));
    invperm_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ix->d_perm.p, n, ix->d_invperm.p);
    KV_CUDA(cudaGetLastError());
    rownorm_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, s>>>(ix->indptr.p, ix->ids.p, ix->tf.p, ix->d_perm.p, n,
                                                                    ix->d_bb64.p, ix->d_univ.p, ix->d_B64.p,
                                                                    ix->d_B32.p, nullptr);
    KV_CUDA(cudaGetLastError());
    chunk_meta_kernel<<<(unsigned)((ix->n_chunks + 255) / 256), 256, 0, s>>>(ix->d_B32.p, n, ix->n_chunks, ix->d_cminB.p);
    KV_CUDA(cudaGetLastError());

    // ---- scan stream + chunk summaries, built per chunk on the host cores ----
    const int T = (int)std::max<int64_t>(1, std::min<int64_t>(host_threads(), ix->n_chunks));
    std::vector<int64_t> chunkptr((size_t)ix->n_chunks + 1, 0), sumptr((size_t)ix->n_chunks + 1, 0);
    std::vector<std::vector<uint32_t>> part((size_t)T), spart((size_t)T);
    std::vector<std::vector<std::pair<unsigned long long, uint32_t>>> povf((size_t)T);
    std::vector<int64_t> chunk_len((size_t)ix->n_chunks, 0), sum_len((size_t)ix->n_chunks, 0), union_len((size_t)ix->n_chunks, 0);
    std::vector<std::vector<std::pair<uint32_t, uint32_t>>> punion((size_t)T);  // chunk unions (fid, max tf), per thread
    parallel_for(ix->n_chunks, T, [&](int t, int64_t c0, int64_t c1) {
      std::vector<std::pair<uint32_t, uint32_t>> u;                       // (fid, tf) of the whole chunk
      std::vector<std::vector<std::pair<uint32_t, uint32_t>>> kept;       // stored entries per row
      std::vector<uint32_t> &out = part[(size_t)t], &sout = spart[(size_t)t];
      auto &ovf = povf[(size_t)t];
      auto emit = [&](std::vector<uint32_t> &dst, unsigned long long key_pos, uint32_t f, uint32_t tfv) {
        if (tfv >= TF_OVF) {
          ovf.emplace_back((key_pos << 32) | f, tfv);
          tfv = TF_OVF;
        }
        dst.push_back((f << 5) | tfv);
      };
      for (int64_t c = c0; c < c1; c++) {
        const int64_t pos0 = c * CHUNK_ROWS, pos1 = std::min<int64_t>(n, pos0 + CHUNK_ROWS);
        kept.resize((size_t)(pos1 - pos0));
        u.clear();
        for (int64_t pos = pos0; pos < pos1; pos++) {
          auto &k = kept[(size_t)(pos - pos0)];
          k.clear();
          const int64_t r = perm[(size_t)pos];
          for (int64_t p = ix->h_indptr[(size_t)r]; p < ix->h_indptr[(size_t)r + 1]; p++) {
            uint32_t f = ix->h_ids[(size_t)p];
            if (!ix->h_univ[f]) k.emplace_back(f, ix->h_tf[(size_t)p]);
          }
          u.insert(u.end(), k.begin(), k.end());
        }
        // longest prefix of stored entries shared by every row of the chunk
        size_t lcp = kept[0].size();
        for (size_t i = 1; i < kept.size() && lcp; i++) {
          size_t m = std::min(lcp, kept[i].size()), j = 0;
          while (j < m && kept[i][j] == kept[0][j]) j++;
          lcp = j;
        }
        const size_t before = out.size();
        for (size_t j = 0; j < lcp; j++) emit(out, OVF_CORE_BASE + (unsigned long long)c, kept[0][j].first, kept[0][j].second);
        out.push_back(CORE_ENTRY);
        for (size_t i = 0; i < kept.size(); i++) {
          const size_t row_begin = out.size();
          for (size_t j = lcp; j < kept[i].size(); j++)
            emit(out, (unsigned long long)(pos0 + (int64_t)i), kept[i][j].first, kept[i][j].second);
          if (out.size() == row_begin) out.push_back(PAD_ENTRY);
          out.back() |= 0x80000000u;
        }
        chunk_len[(size_t)c] = (int64_t)(out.size() - before);
        // union of the chunk's stored features with the max tf (summaries are emitted below)
        std::sort(u.begin(), u.end());
        const size_t ubefore = punion[(size_t)t].size();
        for (size_t i = 0; i < u.size();) {
          size_t j = i;
          while (j + 1 < u.size() && u[j + 1].first == u[i].first) j++;
          punion[(size_t)t].emplace_back(u[i].first, u[j].second);  // sorted: last of the run = max tf
          i = j + 1;
        }
        union_len[(size_t)c] = (int64_t)(punion[(size_t)t].size() - ubefore);
      }
    });
    // "summary-universal" features: present in >= 90 % of the chunk unions.  The bounds assume them present
    // in EVERY chunk with their largest tf (a valid over-estimate), so the summaries need not store them.
    {
      std::unique_ptr<std::atomic<uint32_t>[]> cnt(new std::atomic<uint32_t>[(size_t)Vz]);
      std::unique_ptr<std::atomic<uint32_t>[]> tfm(new std::atomic<uint32_t>[(size_t)Vz]);
      parallel_for(Vz, T, [&](int, int64_t a0, int64_t a1) {
        for (int64_t i = a0; i < a1; i++) { cnt[(size_t)i].store(0, std::memory_order_relaxed); tfm[(size_t)i].store(0, std::memory_order_relaxed); }
      });
      parallel_for(T, T, [&](int, int64_t t0, int64_t t1) {
        for (int64_t t = t0; t < t1; t++)
          for (const auto &e : punion[(size_t)t]) {
            cnt[e.first].fetch_add(1, std::memory_order_relaxed);
            uint32_t cur = tfm[e.first].load(std::memory_order_relaxed);
            while (e.second > cur && !tfm[e.first].compare_exchange_weak(cur, e.second, std::memory_order_relaxed)) {}
          }
      });
      ix->h_su.clear();
      const uint32_t need = (uint32_t)std::max<int64_t>(2, (ix->n_chunks * 9 + 9) / 10);
      for (int64_t f = 0; f < V; f++)
        if (cnt[(size_t)f].load(std::memory_order_relaxed) >= need) ix->h_su.emplace_back((uint32_t)f, tfm[(size_t)f].load(std::memory_order_relaxed));
    }
    parallel_for(ix->n_chunks, T, [&](int t, int64_t c0, int64_t c1) {
      std::vector<uint32_t> &sout = spart[(size_t)t];
      auto &ovf = povf[(size_t)t];
      const auto &src = punion[(size_t)t];
      size_t off = 0;
      for (int64_t c = c0; c < c1; c++) {
        // summary pseudo-row: the union minus the summary-universal features
        const size_t sbefore = sout.size();
        for (size_t i = off; i < off + (size_t)union_len[(size_t)c]; i++) {
          {
            auto it = std::lower_bound(ix->h_su.begin(), ix->h_su.end(), std::make_pair(src[i].first, 0u));
            if (it != ix->h_su.end() && it->first == src[i].first) continue;
          }
          uint32_t f = src[i].first, tfv = src[i].second;
          if (tfv >= TF_OVF) {
            ovf.emplace_back(((unsigned long long)(n + c) << 32) | f, tfv);
            tfv = TF_OVF;
          }
          sout.push_back((f << 5) | tfv);
        }
        off += (size_t)union_len[(size_t)c];
        if (sout.size() == sbefore) sout.push_back(PAD_ENTRY);
        sout.back() |= 0x80000000u;
        sum_len[(size_t)c] = (int64_t)(sout.size() - sbefore);
      }
    });
    for (int64_t c = 0; c < ix->n_chunks; c++) {
      chunkptr[(size_t)c + 1] = chunkptr[(size_t)c] + chunk_len[(size_t)c];
      sumptr[(size_t)c + 1] = sumptr[(size_t)c] + sum_len[(size_t)c];
    }
    ix->stream_len = chunkptr[(size_t)ix->n_chunks];
    ix->sum_len = sumptr[(size_t)ix->n_chunks];
    KV_CUDA(ix->d_stream.ensure(ix->stream_len + 32));
    KV_CUDA(ix->d_sum_stream.ensure(ix->sum_len + 32));
    std::vector<std::pair<unsigned long long, uint32_t>> ovf_all;
    for (int t = 0; t < T; t++) {  // thread t produced chunks [n_chunks*t/T, n_chunks*(t+1)/T)
      const int64_t c0 = ix->n_chunks * t / T;
      if (!part[(size_t)t].empty())
        KV_CUDA(cudaMemcpyAsync(ix->d_stream.p + chunkptr[(size_t)c0], part[(size_t)t].data(), part[(size_t)t].size() * 4,
                                cudaMemcpyHostToDevice, s));
      if (!spart[(size_t)t].empty())
        KV_CUDA(cudaMemcpyAsync(ix->d_sum_stream.p + sumptr[(size_t)c0], spart[(size_t)t].data(), spart[(size_t)t].size() * 4,
                                cudaMemcpyHostToDevice, s));
      ovf_all.insert(ovf_all.end(), povf[(size_t)t].begin(), povf[(size_t)t].end());
    }
    KV_CUDA(cudaMemcpyAsync(ix->d_chunkptr.p, chunkptr.data(), (size_t)(ix->n_chunks + 1) * 8, cudaMemcpyHostToDevice, s));
    KV_CUDA(cudaMemcpyAsync(ix->d_sumptr.p, sumptr.data(), (size_t)(ix->n_chunks + 1) * 8, cudaMemcpyHostToDevice, s));
    // ---- bound-pass layout: the chunk summaries again, in groups of SUM_GROUP with their shared part first ----
    {
      const int64_t n_groups = (ix->n_chunks + SUM_GROUP - 1) / SUM_GROUP;
      std::vector<int64_t> uoff((size_t)ix->n_chunks, 0);  // offset of a chunk's union inside its thread's buffer
      std::vector<int> uthr((size_t)ix->n_chunks, 0);
      for (int t = 0; t < T; t++) {
        int64_t off = 0;
        for (int64_t c = ix->n_chunks * t / T; c < ix->n_chunks * (t + 1) / T; c++) {
          uoff[(size_t)c] = off;
          uthr[(size_t)c] = t;
          off += union_len[(size_t)c];
        }
      }
      const int T2 = (int)std::max<int64_t>(1, std::min<int64_t>(T, n_groups));
      std::vector<std::vector<uint32_t>> gp((size_t)T2);
      std::vector<std::vector<std::pair<unsigned long long, uint32_t>>> govf((size_t)T2);
      std::vector<int64_t> glen((size_t)n_groups, 0), grpptr((size_t)n_groups + 1, 0);
      auto is_su = [&](uint32_t f) {
        auto it = std::lower_bound(ix->h_su.begin(), ix->h_su.end(), std::make_pair(f, 0u));
        return it != ix->h_su.end() && it->first == f;
      };
      parallel_for(n_groups, T2, [&](int t, int64_t a, int64_t b) {
        std::vector<std::pair<uint32_t, uint32_t>> core, tmp;
        auto &out = gp[(size_t)t];
        auto push = [&](unsigned long long key_pos, uint32_t f, uint32_t tfv) {
          if (tfv >= TF_OVF) {
            govf[(size_t)t].emplace_back((key_pos << 32) | f, tfv);
            tfv = TF_OVF;
          }
          out.push_back((f << 5) | tfv);
        };
        for (int64_t g = a; g < b; g++) {
          const int64_t c0 = g * SUM_GROUP, c1 = std::min<int64_t>(ix->n_chunks, c0 + SUM_GROUP);
          auto span = [&](int64_t c) {
            const auto &src = punion[(size_t)uthr[(size_t)c]];
            return std::make_pair(src.begin() + (std::ptrdiff_t)uoff[(size_t)c],
                                  src.begin() + (std::ptrdiff_t)(uoff[(size_t)c] + union_len[(size_t)c]));
          };
          // core = entries (fid, tf) present in every chunk union of the group (sorted ranges -> set_intersection)
          core.assign(span(c0).first, span(c0).second);
          for (int64_t c = c0 + 1; c < c1 && !core.empty(); c++) {
            tmp.clear();
            std::set_intersection(core.begin(), core.end(), span(c).first, span(c).second, std::back_inserter(tmp));
            core.swap(tmp);
          }
          const size_t before = out.size();
          for (auto &e : core)
            if (!is_su(e.first)) push(OVF_GCORE_BASE + (unsigned long long)g, e.first, e.second);
          out.push_back(CORE_ENTRY);
          for (int64_t c = c0; c < c1; c++) {
            const size_t row_begin = out.size();
            auto sp = span(c);
            for (auto it = sp.first; it != sp.second; ++it) {
              if (is_su(it->first) || std::binary_search(core.begin(), core.end(), *it)) continue;
              push((unsigned long long)(n + ix->n_chunks + c), it->first, it->second);
            }
            if (out.size() == row_begin) out.push_back(PAD_ENTRY);
            out.back() |= 0x80000000u;
          }
          glen[(size_t)g] = (int64_t)(out.size() - before);
        }
      });
      for (int64_t g = 0; g < n_groups; g++) grpptr[(size_t)g + 1] = grpptr[(size_t)g] + glen[(size_t)g];
      ix->grp_len = grpptr[(size_t)n_groups];
      KV_CUDA(ix->d_grp_stream.ensure(ix->grp_len + 32));
      KV_CUDA(ix->d_grpptr.ensure(n_groups + 1));
      for (int t = 0; t < T2; t++) {
        const int64_t g0 = n_groups * t / T2;
        if (!gp[(size_t)t].empty())
          KV_CUDA(cudaMemcpyAsync(ix->d_grp_stream.p + grpptr[(size_t)g0], gp[(size_t)t].data(), gp[(size_t)t].size() * 4,
                                  cudaMemcpyHostToDevice, s));
        ovf_all.insert(ovf_all.end(), govf[(size_t)t].begin(), govf[(size_t)t].end());
      }
      KV_CUDA(cudaMemcpyAsync(ix->d_grpptr.p, grpptr.data(), (size_t)(n_groups + 1) * 8, cudaMemcpyHostToDevice, s));
      KV_CUDA(cudaStreamSynchronize(s));  // gp goes out of scope
    }
    std::sort(ovf_all.begin(), ovf_all.end());
    ix->n_ovf = (int)ovf_all.size();
    KV_CUDA(ix->d_ovf_keys.ensure(std::max(1, ix->n_ovf))); KV_CUDA(ix->d_ovf_vals.ensure(std::max(1, ix->n_ovf)));
    std::vector<unsigned long long> ok((size_t)ix->n_ovf);
    std::vector<uint32_t> ov((size_t)ix->n_ovf);
    for (int i = 0; i < ix->n_ovf; i++) { ok[(size_t)i] = ovf_all[(size_t)i].first; ov[(size_t)i] = ovf_all[(size_t)i].second; }
    if (ix->n_ovf) {
      KV_CUDA(cudaMemcpyAsync(ix->d_ovf_keys.p, ok.data(), (size_t)ix->n_ovf * 8, cudaMemcpyHostToDevice, s));
      KV_CUDA(cudaMemcpyAsync(ix->d_ovf_vals.p, ov.data(), (size_t)ix->n_ovf * 4, cudaMemcpyHostToDevice, s));
    }
    KV_CUDA(cudaStreamSynchronize(s));  // the staging vectors go out of scope
  } else {
    KV_CUDA(ix->d_stream.ensure(32));
    KV_CUDA(ix->d_sum_stream.ensure(32));
    KV_CUDA(ix->d_grp_stream.ensure(32));
    KV_CUDA(ix->d_grpptr.ensure(1));
    KV_CUDA(cudaMemsetAsync(ix->d_grpptr.p, 0, sizeof(int64_t), s));
    KV_CUDA(cudaMemsetAsync(ix->d_chunkptr.p, 0, sizeof(int64_t), s));
    KV_CUDA(cudaMemsetAsync(ix->d_sumptr.p, 0, sizeof(int64_t), s));
    KV_CUDA(cudaStreamSynchronize(s));
  }
  ix->V = V;
  ix->finalized = true;
  ix->batch_valid = false;
  ix->layout_valid = n > 0;
  ix->layout_rows = n;
  ix->layout_univ = ix->h_univ;
  ix->last_finalize_kind = 1;
  return KV_OK;
}

int kv_index_last_finalize_kind(const kv_index *ix) { return ix ? ix->last_finalize_kind : 0; }

// caller holds ix->mu; scores land in ix->d_scores and, when out_scores != NULL, on the host
static int score_impl(kv_index *ix, const uint32_t *q_ids, const uint32_t *q_tf, int64_t q_nnz, double q_oov_tf2,
                      double *out_scores) {
  if (!ix->finalized) return kv_fail(KV_ERR_STATE, "kv_score: index not finalized");
  if (ix->n_rows == 0) return KV_OK;
  if (ix->nnz == 0 && q_nnz == 0 && q_oov_tf2 == 0.0)
    return kv_fail(KV_ERR_EMPTY_VOCAB, "empty vocabulary; perhaps the documents only contain stop words");
  KV_CUDA(cudaSetDevice(ix->device));
  QueryPrep qp;
  prep_query(ix, q_ids, q_tf, q_nnz, q_oov_tf2, qp);
  int log_h = 6;
  while ((1 << log_h) < 2 * (int)qp.fid.size() + 2) log_h++;
  if (log_h > 24) return kv_fail(KV_ERR_INVALID, "kv_score: query has too many distinct features");
  const int H = 1 << log_h;
  const size_t tab_bytes = (size_t)H * (8 + 8 + 4);
  KV_CUDA(ix->h_qtab.ensure((int64_t)tab_bytes));
  KV_CUDA(ix->d_qtab.ensure((int64_t)tab_bytes));
  double *tw = (double *)ix->h_qtab.p, *tdd = tw + H;
  uint32_t *tk = (uint32_t *)(tdd + H);
  for (int i = 0; i < H; i++) { tk[i] = KEY_EMPTY; tw[i] = 0; tdd[i] = 0; }
  for (size_t i = 0; i < qp.fid.size(); i++) {
    uint32_t t = qp.fid[i];
    uint32_t h = (t * 0x9E3779B1u) >> (32 - log_h);
    while (tk[h] != KEY_EMPTY) h = (h + 1) & (H - 1);
    double a, d;
    idf_host(ix->n_total, ix->h_df[t], a, d, ix->jaccard, ix->corpus_fit);
    tk[h] = t;
    tw[h] = (double)qp.tfq[i] * a;
    tdd[h] = d;
  }
  cudaStream_t s = ix->stream;
  KV_CUDA(cudaMemcpyAsync(ix->d_qtab.p, ix->h_qtab.p, tab_bytes, cudaMemcpyHostToDevice, s));
  KV_CUDA(ix->d_scores.ensure(ix->n_rows));
  ScoreParams P;
  P.stream = ix->d_stream.p; P.chunkptr = ix->d_chunkptr.p; P.perm = ix->d_perm.p;
  P.n_chunks = ix->n_chunks; P.n_rows = ix->n_rows;
  P.B64 = ix->d_B64.p; P.ovf_keys = ix->d_ovf_keys.p; P.ovf_vals = ix->d_ovf_vals.p; P.n_ovf = ix->n_ovf;
  P.qw = (const double *)ix->d_qtab.p; P.qd = P.qw + H; P.qkeys = (const uint32_t *)(P.qd + H);
  P.log_h = log_h; P.table_in_smem = tab_bytes <= 40 * 1024;
  P.nq = qp.nq; P.dotU = qp.dotU; P.corrU = qp.corrU; P.jaccard = ix->jaccard; P.out = ix->d_scores.p;
  int blocks = (int)std::min<int64_t>((ix->n_chunks + 7) / 8, (int64_t)ix->sm_count * 8);
  if (blocks < 1) blocks = 1;
  KV_CUDA(cudaEventRecord(ix->ev[1], s));
  tfidf_score_kernel<<<blocks, 256, P.table_in_smem ? tab_bytes : 0, s>>>(P);
  KV_CUDA(cudaGetLastError());
  KV_CUDA(cudaEventRecord(ix->ev[2], s));
  if (out_scores)
    KV_CUDA(cudaMemcpyAsync(out_scores, ix->d_scores.p, (size_t)ix->n_rows * sizeof(double), cudaMemcpyDeviceToHost, s));
  KV_CUDA(cudaStreamSynchronize(s));
  cudaEventElapsedTime(&ix->last_score_ms, ix->ev[1], ix->ev[2]);
  return KV_OK;
}

int kv_score(kv_index *ix, const uint32_t *q_ids, const uint32_t *q_tf, int64_t q_nnz, double q_oov_tf2,
             double *out_scores) {
  if (!ix || q_nnz < 0 || (q_nnz > 0 && (!q_ids || !q_tf)))
    return kv_fail(KV_ERR_INVALID, "kv_score: bad arguments");
  std::lock_guard<std::mutex> g(ix->mu);
  if (ix->finalized && ix->n_rows > 0 && !out_scores) return kv_fail(KV_ERR_INVALID, "kv_score: out_scores is NULL");
  return score_impl(ix, q_ids, q_tf, q_nnz, q_oov_tf2, out_scores);
}

// ---- batched top-k, in two halves: prepare_batch (host work + H2D) and run_batch (device only) ----
static int prepare_batch(kv_index *ix, const int64_t *q_indptr, const uint32_t *q_ids, const uint32_t *q_tf,
                         const double *q_oov, int64_t n_q) {
  if (!ix->finalized) return kv_fail(KV_ERR_STATE, "kv_topk: index not finalized");
  if (n_q >= (1LL << 31)) return kv_fail(KV_ERR_INVALID, "kv_topk: too many queries in one call");
  KV_CUDA(cudaSetDevice(ix->device));
  cudaStream_t s = ix->stream;
  ix->batch_valid = false;
  ix->has_excl = false;
  ix->irr_q.clear(); ix->irr_indptr.assign(1, 0); ix->irr_ids.clear(); ix->irr_tf.clear(); ix->irr_oov.clear();
  for (int64_t q = 0; q < n_q; q++)
    if (q_indptr[q + 1] < q_indptr[q] || (q_indptr[q + 1] > q_indptr[q] && (!q_ids || !q_tf)))
      return kv_fail(KV_ERR_INVALID, "kv_topk: bad query CSR");

  // ---- host: per-query constants (float64), text order of the queries, tile tables ----
  KV_CUDA(ix->h_qconst.ensure(6 * n_q));
  KV_CUDA(ix->h_qperm.ensure(2 * n_q));
  std::vector<QueryPrep> qp((size_t)n_q);
  const int T = host_threads();
  parallel_for(n_q, n_q >= 2048 ? T : 1, [&](int, int64_t a, int64_t b) {
    for (int64_t q = a; q < b; q++)
      prep_query(ix, q_ids + q_indptr[q], q_tf + q_indptr[q], q_indptr[q + 1] - q_indptr[q], q_oov ? q_oov[q] : 0.0,
                 qp[(size_t)q]);
  });
  // queries with similar text share a tile: smaller feature tables, and tile-wide pruning works
  std::vector<int> order((size_t)n_q);
  for (int64_t q = 0; q < n_q; q++) order[(size_t)q] = (int)q;
  // first key: class of the query norm (queries with much out-of-corpus mass have low scores and low
  // k-th-score thresholds; mixing them with strong queries would block the tile-wide pruning)
  std::vector<short> qcls((size_t)n_q);
  for (int64_t q = 0; q < n_q; q++)
    qcls[(size_t)q] = qp[(size_t)q].nq > 0 ? (short)std::floor(std::log2(qp[(size_t)q].nq) * 2.0) : (short)-1000;
  stable_sort_indices(order, [&](int a, int b) {
    if (qcls[(size_t)a] != qcls[(size_t)b]) return qcls[(size_t)a] < qcls[(size_t)b];
    return cmp_seq(q_ids + q_indptr[a], q_indptr[a + 1] - q_indptr[a], q_ids + q_indptr[b], q_indptr[b + 1] - q_indptr[b]) < 0;
  }, T);  // == std::stable_sort, on all host threads
  float *c_nq = ix->h_qconst.p, *c_dotU = c_nq + n_q, *c_corrU = c_dotU + n_q, *c_ninf = c_corrU + n_q;
  float *c_dotS = c_ninf + n_q, *c_corrS = c_dotS + n_q;
  int *qperm = ix->h_qperm.p, *null_list = qperm + n_q;
  int64_t n_null = 0;
  std::vector<char> skip((size_t)n_q, 0);  // by sorted slot: not part of any tile table
  for (int64_t i = 0; i < n_q; i++) {
    const int64_t q = order[(size_t)i];
    const QueryPrep &p = qp[(size_t)q];
    qperm[i] = (int)q;
    c_dotU[i] = (float)p.dotU;
    c_corrU[i] = (float)p.corrU;
    c_ninf[i] = -INFINITY;
    c_dotS[i] = (float)(p.dotS * (1.0 + 1e-6));  // bounds may only err upwards
    c_corrS[i] = (float)p.corrS;
    int nx = 0;
    for (uint32_t f : p.tfq) nx += f > 1;
    const bool null_q = p.nq <= 0.0 || (p.fid.empty() && p.dotU == 0.0);  // every score is 0
    const bool irregular = (int)p.fid.size() > TILE_MAX_FEATURES || nx > TXCAP;
    c_nq[i] = (null_q || irregular) ? 0.f : (float)p.nq;  // nq == 0 switches the lane off in the kernel
    if (null_q) {
      null_list[n_null++] = (int)q;
      skip[(size_t)i] = 1;
    } else if (irregular) {  // too many features for a tile: full float64 scan + selection instead
      const int64_t a = q_indptr[q], b = q_indptr[q + 1];
      ix->irr_q.push_back(q);
      ix->irr_ids.insert(ix->irr_ids.end(), q_ids + a, q_ids + b);
      ix->irr_tf.insert(ix->irr_tf.end(), q_tf + a, q_tf + b);
      ix->irr_indptr.push_back((int64_t)ix->irr_ids.size());
      ix->irr_oov.push_back(q_oov ? q_oov[q] : 0.0);
      skip[(size_t)i] = 1;
    }
  }
  // tiles: consecutive sorted queries, closed when 128 queries are in or the feature table is full (tile_builder.cuh:
  // built on all host threads when no table cap is hit, else by the sequential rule -- identical bytes either way)
  std::vector<TileDesc> tiles;
  std::vector<unsigned char> tables;
  {
    const TileCtx cx{ix->n_total, ix->h_df.data(), ix->jaccard, ix->corpus_fit};
    if (getenv("KAKVEDA_B200_SERIAL_TILES") || !build_tiles_parallel(cx, qp, order, skip, T, tiles, tables))
      build_tiles_serial(cx, qp, order, skip, tiles, tables);
  }
  const int64_t n_tiles = (int64_t)tiles.size();
  KV_CUDA(ix->h_tables.ensure(n_tiles * (int64_t)Tile::table_bytes));
  KV_CUDA(ix->h_tiles.ensure(n_tiles));
  memcpy(ix->h_tables.p, tables.data(), (size_t)n_tiles * Tile::table_bytes);
  memcpy(ix->h_tiles.p, tiles.data(), (size_t)n_tiles * sizeof(TileDesc));

  KV_CUDA(ix->d_tables.ensure(n_tiles * (int64_t)Tile::table_bytes));
  KV_CUDA(ix->d_tiles.ensure(n_tiles));
  KV_CUDA(ix->d_qconst.ensure(6 * n_q));
  KV_CUDA(ix->d_qperm.ensure(2 * n_q));
  KV_CUDA(cudaEventRecord(ix->ev[0], s));
  KV_CUDA(cudaMemcpyAsync(ix->d_tables.p, ix->h_tables.p, (size_t)n_tiles * Tile::table_bytes, cudaMemcpyHostToDevice, s));
  KV_CUDA(cudaMemcpyAsync(ix->d_tiles.p, ix->h_tiles.p, (size_t)n_tiles * sizeof(TileDesc), cudaMemcpyHostToDevice, s));
  KV_CUDA(cudaMemcpyAsync(ix->d_qconst.p, ix->h_qconst.p, (size_t)6 * n_q * sizeof(float), cudaMemcpyHostToDevice, s));
  KV_CUDA(cudaMemcpyAsync(ix->d_qperm.p, ix->h_qperm.p, (size_t)2 * n_q * sizeof(int), cudaMemcpyHostToDevice, s));
  KV_CUDA(cudaEventRecord(ix->ev[1], s));
  KV_CUDA(cudaStreamSynchronize(s));  // the pinned staging buffers may be rewritten by the next call
  ix->batch_q = n_q;
  ix->batch_tiles = n_tiles;
  ix->batch_null = n_null;
  ix->batch_h2d_bytes = n_tiles * (int64_t)(Tile::table_bytes + sizeof(TileDesc)) + 6 * n_q * (int64_t)sizeof(float) +
                        2 * n_q * (int64_t)sizeof(int);
  ix->batch_valid = true;
  cudaEventElapsedTime(&ix->last_ms[0], ix->ev[0], ix->ev[1]);
  return KV_OK;
}

// Device-only half: scan + merge (+ fallback scans for irregular queries) of the uploaded batch.
static int run_batch(kv_index *ix, int k, float *d_out_s, long long *d_out_r) {
  if (!ix->batch_valid) return kv_fail(KV_ERR_STATE, "kv_topk_resident: no query batch uploaded");
  if (k < 1 || k > 32) return kv_fail(KV_ERR_INVALID, "kv_topk: k must be 1..32");
  KV_CUDA(cudaSetDevice(ix->device));
  cudaStream_t s = ix->stream;
  const int64_t n_q = ix->batch_q, n_tiles = ix->batch_tiles;
  // launch geometry: tiles x row-splits.  Without pruning every CTA streams its whole row range, so
  // aim at >= 8 waves of resident CTAs; with pruning a CTA owns a tile's whole row range unless
  // there are too few tiles to fill the GPU.
  const char *env = getenv("KAKVEDA_B200_NO_PRUNE");
  const int prune = (env && env[0] == '1') ? 0 : (ix->n_chunks >= 64 ? 1 : 0);
  const int ctas_per_sm = 3;
  int64_t want = (int64_t)ix->sm_count * ctas_per_sm * 8;
  int64_t n_splits = (want + n_tiles - 1) / n_tiles;
  n_splits = std::max<int64_t>(1, std::min<int64_t>(n_splits, std::max<int64_t>(1, ix->n_chunks / (2 * SUM_GROUP))));
  n_splits = std::min<int64_t>(n_splits, 2048);
  ix->last_tiles = n_tiles; ix->last_splits = n_splits; ix->last_ctas = n_tiles * n_splits;
  if (n_q > ix->d_gthr.cap && (ix->gthr_exported || ix->n_peers))
    return kv_fail(KV_ERR_STATE, "kv_topk: the query batch outgrew the threshold array shared with the peer GPUs; exchange it again "
                                 "(kv_index_thresholds_export / kv_index_thresholds_peers)");
  KV_CUDA(ix->d_gthr.ensure(n_q));
  const int n_peers = (ix->n_peers > 0 && n_q <= ix->peer_cap) ? ix->n_peers : 0;
  KV_CUDA(ix->d_part_s.ensure(n_splits * n_q * k));
  KV_CUDA(ix->d_part_r.ensure(n_splits * n_q * k));
  KV_CUDA(ix->d_stats.ensure(8));
  if (prune) KV_CUDA(ix->d_ubuf.ensure(n_tiles * ix->n_chunks));

  KV_CUDA(cudaEventRecord(ix->ev[1], s));
  if (ix->n_rows > 0) {
    // global lower bounds of the k-th score start at -inf (staged as the 4th constants column)
    KV_CUDA(cudaMemcpyAsync(ix->d_gthr.p, ix->d_qconst.p + 3 * n_q, (size_t)n_q * sizeof(float), cudaMemcpyDeviceToDevice, s));
    KV_CUDA(cudaMemsetAsync(ix->d_stats.p, 0, 8 * sizeof(unsigned long long), s));
    TopkParams P;
    P.stream = ix->d_stream.p; P.chunkptr = ix->d_chunkptr.p; P.sum_stream = ix->d_sum_stream.p; P.sumptr = ix->d_sumptr.p;
    P.chunk_minB = ix->d_cminB.p; P.perm = ix->d_perm.p;
    P.grp_stream = ix->d_grp_stream.p; P.grpptr = ix->d_grpptr.p;
    P.n_chunks = ix->n_chunks; P.n_rows = ix->n_rows;
    P.row_base = ix->row_base; P.B32 = ix->d_B32.p; P.ovf_keys = ix->d_ovf_keys.p; P.ovf_vals = ix->d_ovf_vals.p;
    P.n_ovf = ix->n_ovf; P.tables = ix->d_tables.p; P.tiles = ix->d_tiles.p;
    P.q_nq = ix->d_qconst.p; P.q_dotU = P.q_nq + n_q; P.q_corrU = P.q_dotU + n_q;
    P.q_dotS = P.q_nq + 4 * n_q; P.q_corrS = P.q_nq + 5 * n_q;
    P.q_excl = ix->has_excl ? ix->d_excl_sorted.p : nullptr;
    P.gthr = ix->d_gthr.p; P.ubuf = ix->d_ubuf.p; P.stats = ix->d_stats.p;
    P.n_q = n_q; P.k = k; P.n_splits = (int)n_splits; P.prune = prune; P.jaccard = ix->jaccard;
    for (int i = 0; i < 7; i++) P.peer_gthr[i] = i < n_peers ? ix->peer_gthr[i] : nullptr;
    P.n_peers = n_peers;
    P.share = (n_splits > 1 || n_peers > 0) ? 1 : 0;
    P.part_scores = ix->d_part_s.p; P.part_rows = ix->d_part_r.p;
    const size_t smem = Tile::smem_bytes(k);
    static bool attr_set[64] = {false};
    if (!attr_set[ix->device & 63]) {
      KV_CUDA(cudaFuncSetAttribute(tfidf_topk_kernel<TG, TLOGH, TXCAP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)Tile::smem_bytes(32)));
      attr_set[ix->device & 63] = true;
    }
    dim3 grid((unsigned)n_tiles, (unsigned)n_splits);
    tfidf_topk_kernel<TG, TLOGH, TXCAP><<<grid, 256, smem, s>>>(P);
    KV_CUDA(cudaGetLastError());
    KV_CUDA(cudaEventRecord(ix->ev[2], s));
    merge_topk_kernel<<<(unsigned)((n_q * 32 + 255) / 256), 256, 0, s>>>(ix->d_part_s.p, ix->d_part_r.p, (int)n_splits,
                                                                         n_q, k, ix->d_qperm.p, d_out_s, d_out_r);
    KV_CUDA(cudaGetLastError());
    if (ix->batch_null) {
      fill_null_kernel<<<(unsigned)((ix->batch_null * k + 255) / 256), 256, 0, s>>>(ix->d_qperm.p + n_q, (int)ix->batch_null, k,
                                                                                    ix->n_rows, ix->row_base,
                                                                                    ix->has_excl ? ix->d_excl_orig.p : nullptr, d_out_s, d_out_r);
      KV_CUDA(cudaGetLastError());
    }
    for (size_t i = 0; i < ix->irr_q.size(); i++) {
      const int64_t q = ix->irr_q[i], a = ix->irr_indptr[i], b = ix->irr_indptr[i + 1];
      int rc = score_impl(ix, ix->irr_ids.data() + a, ix->irr_tf.data() + a, b - a, ix->irr_oov[i], nullptr);
      if (rc != KV_OK) return rc;
      select_topk_kernel<<<1, 1024, 0, s>>>(ix->d_scores.p, ix->n_rows, ix->row_base, k,
                                            ix->has_excl ? (int64_t)ix->h_excl_orig[(size_t)q] : -1, d_out_s + q * k, d_out_r + q * k);
      KV_CUDA(cudaGetLastError());
    }
    KV_CUDA(cudaMemcpyAsync(ix->last_stats, ix->d_stats.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  } else {
    KV_CUDA(cudaEventRecord(ix->ev[2], s));
    std::vector<float> es((size_t)(n_q * k), -INFINITY);
    std::vector<long long> er((size_t)(n_q * k), -1);
    KV_CUDA(cudaMemcpyAsync(d_out_s, es.data(), es.size() * 4, cudaMemcpyHostToDevice, s));
    KV_CUDA(cudaMemcpyAsync(d_out_r, er.data(), er.size() * 8, cudaMemcpyHostToDevice, s));
    KV_CUDA(cudaStreamSynchronize(s));
  }
  KV_CUDA(cudaEventRecord(ix->ev[3], s));
  return KV_OK;
}

static int topk_impl(kv_index *ix, const int64_t *q_indptr, const uint32_t *q_ids, const uint32_t *q_tf,
                     const double *q_oov, int64_t n_q, int k, float *h_scores, int64_t *h_rows, void *d_scores_out,
                     void *d_rows_out) {
  if (!ix || n_q < 0 || k < 1 || k > 32 || (n_q > 0 && !q_indptr))
    return kv_fail(KV_ERR_INVALID, "kv_topk: bad arguments (k must be 1..32)");
  std::lock_guard<std::mutex> g(ix->mu);
  if (!ix->finalized) return kv_fail(KV_ERR_STATE, "kv_topk: index not finalized");
  if (n_q == 0) return KV_OK;
  int rc = prepare_batch(ix, q_indptr, q_ids, q_tf, q_oov, n_q);
  if (rc != KV_OK) return rc;
  cudaStream_t s = ix->stream;
  float *d_out_s = (float *)d_scores_out;
  long long *d_out_r = (long long *)d_rows_out;
  if (!d_out_s) {
    KV_CUDA(ix->d_out_s.ensure(n_q * k)); KV_CUDA(ix->d_out_r.ensure(n_q * k));
    d_out_s = ix->d_out_s.p; d_out_r = ix->d_out_r.p;
  }
  rc = run_batch(ix, k, d_out_s, d_out_r);
  if (rc != KV_OK) return rc;
  if (h_scores) {
    KV_CUDA(ix->h_out_s.ensure(n_q * k)); KV_CUDA(ix->h_out_r.ensure(n_q * k));
    KV_CUDA(cudaMemcpyAsync(ix->h_out_s.p, d_out_s, (size_t)n_q * k * sizeof(float), cudaMemcpyDeviceToHost, s));
    KV_CUDA(cudaMemcpyAsync(ix->h_out_r.p, d_out_r, (size_t)n_q * k * sizeof(long long), cudaMemcpyDeviceToHost, s));
  }
  KV_CUDA(cudaEventRecord(ix->ev[4], s));
  KV_CUDA(cudaStreamSynchronize(s));
  if (h_scores) {
    memcpy(h_scores, ix->h_out_s.p, (size_t)n_q * k * sizeof(float));
    memcpy(h_rows, ix->h_out_r.p, (size_t)n_q * k * sizeof(int64_t));
  }
  for (int i = 1; i < 4; i++) cudaEventElapsedTime(&ix->last_ms[i], ix->ev[i], ix->ev[i + 1]);
  return KV_OK;
}

int kv_query_upload(kv_index *ix, const int64_t *q_indptr, const uint32_t *q_ids, const uint32_t *q_tf,
                    const double *q_oov_tf2, int64_t n_q) {
  if (!ix || n_q < 1 || !q_indptr) return kv_fail(KV_ERR_INVALID, "kv_query_upload: bad arguments");
  std::lock_guard<std::mutex> g(ix->mu);
  return prepare_batch(ix, q_indptr, q_ids, q_tf, q_oov_tf2, n_q);
}

static int set_exclusions_locked(kv_index *ix, const int64_t *exclude_rows, int64_t n_q);

int kv_query_set_exclusions(kv_index *ix, const int64_t *exclude_rows, int64_t n_q) {
  if (!ix) return kv_fail(KV_ERR_INVALID, "kv_query_set_exclusions: NULL handle");
  std::lock_guard<std::mutex> g(ix->mu);
  return set_exclusions_locked(ix, exclude_rows, n_q);
}

int kv_selfjoin_upload(kv_index *ix, int64_t q_begin, int64_t q_end) {
  if (!ix || q_begin < 0 || q_end <= q_begin) return kv_fail(KV_ERR_INVALID, "kv_selfjoin_upload: bad row range");
  std::lock_guard<std::mutex> g(ix->mu);
  if (!ix->finalized) return kv_fail(KV_ERR_STATE, "kv_selfjoin_upload: index not finalized");
  if (q_end > ix->n_rows) return kv_fail(KV_ERR_INVALID, "kv_selfjoin_upload: row range outside the index");
  const int64_t n = q_end - q_begin, base = ix->h_indptr[(size_t)q_begin];
  std::vector<int64_t> ip((size_t)n + 1), ex((size_t)n);
  for (int64_t i = 0; i <= n; i++) ip[(size_t)i] = ix->h_indptr[(size_t)(q_begin + i)] - base;
  std::vector<uint32_t> tf32((size_t)ip[(size_t)n]);
  for (size_t i = 0; i < tf32.size(); i++) tf32[i] = ix->h_tf[(size_t)base + i];
  int rc = prepare_batch(ix, ip.data(), ix->h_ids.data() + base, tf32.data(), nullptr, n);
  if (rc != KV_OK) return rc;
  for (int64_t i = 0; i < n; i++) ex[(size_t)i] = ix->row_base + q_begin + i;
  return set_exclusions_locked(ix, ex.data(), n);
}

int kv_rescore_pairs(kv_index *ix, const int64_t *q_indptr, const uint32_t *q_ids, const uint32_t *q_tf,
                     const double *q_oov_tf2, int64_t n_q, int k, const int64_t *rows, double *out_scores) {
  if (!ix || n_q < 0 || k < 1 || (n_q > 0 && (!q_indptr || !rows || !out_scores)))
    return kv_fail(KV_ERR_INVALID, "kv_rescore_pairs: bad arguments");
  std::lock_guard<std::mutex> g(ix->mu);
  if (!ix->finalized) return kv_fail(KV_ERR_STATE, "kv_rescore_pairs: index not finalized");
  if (n_q == 0) return KV_OK;
  const int64_t nnz = q_indptr[n_q] - q_indptr[0];
  if (nnz < 0 || (nnz > 0 && (!q_ids || !q_tf))) return kv_fail(KV_ERR_INVALID, "kv_rescore_pairs: bad query CSR");
  if (ix->n_rows == 0) {
    for (int64_t i = 0; i < n_q * k; i++) out_scores[i] = -INFINITY;
    return KV_OK;
  }
  KV_CUDA(cudaSetDevice(ix->device));
  cudaStream_t s = ix->stream;
  // |q|^2 exactly as K1a gets it (prep_query, float64)
  std::vector<double> cst((size_t)n_q);
  std::vector<int64_t> ip((size_t)n_q + 1);
  for (int64_t q = 0; q <= n_q; q++) ip[(size_t)q] = q_indptr[q] - q_indptr[0];
  parallel_for(n_q, n_q >= 2048 ? host_threads() : 1, [&](int, int64_t a, int64_t b) {
    QueryPrep qp;
    for (int64_t q = a; q < b; q++) {
      prep_query(ix, q_ids + q_indptr[q], q_tf + q_indptr[q], q_indptr[q + 1] - q_indptr[q], q_oov_tf2 ? q_oov_tf2[q] : 0.0, qp);
      cst[(size_t)q] = qp.nq;
    }
  });
  KV_CUDA(ix->d_rq_indptr.ensure(n_q + 1)); KV_CUDA(ix->d_rq_ids.ensure(std::max<int64_t>(nnz, 1)));
  KV_CUDA(ix->d_rq_tf.ensure(std::max<int64_t>(nnz, 1))); KV_CUDA(ix->d_rq_const.ensure(n_q));
  KV_CUDA(ix->d_rq_out.ensure(n_q * k)); KV_CUDA(ix->d_rq_rows.ensure(n_q * k));
  KV_CUDA(cudaMemcpyAsync(ix->d_rq_indptr.p, ip.data(), (size_t)(n_q + 1) * 8, cudaMemcpyHostToDevice, s));
  if (nnz) {
    KV_CUDA(cudaMemcpyAsync(ix->d_rq_ids.p, q_ids + q_indptr[0], (size_t)nnz * 4, cudaMemcpyHostToDevice, s));
    KV_CUDA(cudaMemcpyAsync(ix->d_rq_tf.p, q_tf + q_indptr[0], (size_t)nnz * 4, cudaMemcpyHostToDevice, s));
  }
  KV_CUDA(cudaMemcpyAsync(ix->d_rq_const.p, cst.data(), (size_t)n_q * 8, cudaMemcpyHostToDevice, s));
  KV_CUDA(cudaMemcpyAsync(ix->d_rq_rows.p, rows, (size_t)n_q * k * 8, cudaMemcpyHostToDevice, s));
  RescoreParams P;
  P.indptr = ix->indptr.p; P.ids = ix->ids.p; P.tf = ix->tf.p;
  P.a64 = ix->d_a64.p; P.d64 = ix->d_d64.p; P.B64 = ix->d_B64.p; P.invperm = ix->d_invperm.p;
  P.q_indptr = ix->d_rq_indptr.p; P.q_ids = ix->d_rq_ids.p; P.q_tf = ix->d_rq_tf.p;
  P.q_nq = ix->d_rq_const.p;
  P.rows = ix->d_rq_rows.p; P.n_q = n_q; P.n_rows = ix->n_rows; P.row_base = ix->row_base; P.V = ix->V;
  P.k = k; P.jaccard = ix->jaccard; P.out = ix->d_rq_out.p;
  rescore_kernel<<<(unsigned)((n_q * k * 32 + 255) / 256), 256, 0, s>>>(P);
  KV_CUDA(cudaGetLastError());
  KV_CUDA(cudaMemcpyAsync(out_scores, ix->d_rq_out.p, (size_t)n_q * k * 8, cudaMemcpyDeviceToHost, s));
  KV_CUDA(cudaStreamSynchronize(s));
  return KV_OK;
}

static int set_exclusions_locked(kv_index *ix, const int64_t *exclude_rows, int64_t n_q) {
  if (!ix->batch_valid) return kv_fail(KV_ERR_STATE, "kv_query_set_exclusions: no query batch uploaded");
  if (!exclude_rows) { ix->has_excl = false; return KV_OK; }
  if (n_q != ix->batch_q) return kv_fail(KV_ERR_INVALID, "kv_query_set_exclusions: %lld entries for a batch of %lld queries",
                                         (long long)n_q, (long long)ix->batch_q);
  KV_CUDA(cudaSetDevice(ix->device));
  // global row ids -> local original rows of this shard (-1: none, or the row lives on another shard)
  ix->h_excl_orig.assign((size_t)n_q, -1);
  for (int64_t q = 0; q < n_q; q++) {
    const int64_t r = exclude_rows[q] - ix->row_base;
    if (exclude_rows[q] >= 0 && r >= 0 && r < ix->n_rows) ix->h_excl_orig[(size_t)q] = (int)r;
  }
  std::vector<int> sorted((size_t)n_q);
  for (int64_t i = 0; i < n_q; i++) sorted[(size_t)i] = ix->h_excl_orig[(size_t)ix->h_qperm.p[i]];
  KV_CUDA(ix->d_excl_sorted.ensure(n_q)); KV_CUDA(ix->d_excl_orig.ensure(n_q));
  KV_CUDA(cudaMemcpyAsync(ix->d_excl_sorted.p, sorted.data(), (size_t)n_q * sizeof(int), cudaMemcpyHostToDevice, ix->stream));
  KV_CUDA(cudaMemcpyAsync(ix->d_excl_orig.p, ix->h_excl_orig.data(), (size_t)n_q * sizeof(int), cudaMemcpyHostToDevice, ix->stream));
  KV_CUDA(cudaStreamSynchronize(ix->stream));
  ix->has_excl = true;
  return KV_OK;
}

int kv_topk_resident_host(kv_index *ix, int k, float *out_scores, int64_t *out_rows) {
  if (!ix || !out_scores || !out_rows) return kv_fail(KV_ERR_INVALID, "kv_topk_resident_host: bad arguments");
  std::lock_guard<std::mutex> g(ix->mu);
  if (!ix->batch_valid) return kv_fail(KV_ERR_STATE, "kv_topk_resident_host: no query batch uploaded");
  if (k < 1 || k > 32) return kv_fail(KV_ERR_INVALID, "kv_topk: k must be 1..32");
  const int64_t n_q = ix->batch_q;
  KV_CUDA(cudaSetDevice(ix->device));
  KV_CUDA(ix->d_out_s.ensure(n_q * k)); KV_CUDA(ix->d_out_r.ensure(n_q * k));
  int rc = run_batch(ix, k, ix->d_out_s.p, ix->d_out_r.p);
  if (rc != KV_OK) return rc;
  KV_CUDA(cudaMemcpyAsync(out_scores, ix->d_out_s.p, (size_t)n_q * k * sizeof(float), cudaMemcpyDeviceToHost, ix->stream));
  KV_CUDA(cudaMemcpyAsync(out_rows, ix->d_out_r.p, (size_t)n_q * k * sizeof(long long), cudaMemcpyDeviceToHost, ix->stream));
  KV_CUDA(cudaEventRecord(ix->ev[4], ix->stream));
  KV_CUDA(cudaStreamSynchronize(ix->stream));
  for (int i = 1; i < 4; i++) cudaEventElapsedTime(&ix->last_ms[i], ix->ev[i], ix->ev[i + 1]);
  return KV_OK;
}

int kv_index_thresholds_export(kv_index *ix, int64_t capacity, void *handle_out) {
  if (!ix || capacity < 1 || !handle_out) return kv_fail(KV_ERR_INVALID, "kv_index_thresholds_export: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the C ABI documents a 64-byte handle");
  std::lock_guard<std::mutex> g(ix->mu);
  KV_CUDA(cudaSetDevice(ix->device));
  if (ix->n_peers) return kv_fail(KV_ERR_STATE, "kv_index_thresholds_export: clear the peer mapping first (n_peers = 0)");
  KV_CUDA(ix->d_gthr.ensure(capacity));
  cudaIpcMemHandle_t h;
  KV_CUDA(cudaIpcGetMemHandle(&h, ix->d_gthr.p));
  memcpy(handle_out, &h, sizeof(h));
  ix->gthr_exported = true;
  return KV_OK;
}

int kv_index_thresholds_peers(kv_index *ix, const void *handles, int n_peers, int64_t capacity) {
  if (!ix || n_peers < 0 || n_peers > 7 || (n_peers > 0 && (!handles || capacity < 1)))
    return kv_fail(KV_ERR_INVALID, "kv_index_thresholds_peers: bad arguments (at most 7 peers)");
  std::lock_guard<std::mutex> g(ix->mu);
  KV_CUDA(cudaSetDevice(ix->device));
  KV_CUDA(cudaStreamSynchronize(ix->stream));
  close_peers(ix);
  if (n_peers == 0) { ix->gthr_exported = false; return KV_OK; }
  for (int i = 0; i < n_peers; i++) {
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)handles + (size_t)i * sizeof(h), sizeof(h));
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      ix->n_peers = i;
      close_peers(ix);
      return kv_fail(KV_ERR_CUDA, "kv_index_thresholds_peers: cudaIpcOpenMemHandle failed for peer %d: %s", i, cudaGetErrorString(e));
    }
    ix->peer_gthr[i] = (int *)p;
  }
  ix->n_peers = n_peers;
  ix->peer_cap = capacity;
  return KV_OK;
}

int kv_topk_resident(kv_index *ix, int k, void *d_scores, void *d_rows) {
  if (!ix || !d_scores || !d_rows) return kv_fail(KV_ERR_INVALID, "kv_topk_resident: bad arguments");
  std::lock_guard<std::mutex> g(ix->mu);
  int rc = run_batch(ix, k, (float *)d_scores, (long long *)d_rows);
  if (rc != KV_OK) return rc;
  KV_CUDA(cudaEventRecord(ix->ev[4], ix->stream));
  KV_CUDA(cudaStreamSynchronize(ix->stream));
  for (int i = 1; i < 4; i++) cudaEventElapsedTime(&ix->last_ms[i], ix->ev[i], ix->ev[i + 1]);
  return KV_OK;
}

int kv_topk(kv_index *ix, const int64_t *q_indptr, const uint32_t *q_ids, const uint32_t *q_tf,
            const double *q_oov_tf2, int64_t n_q, int k, float *out_scores, int64_t *out_rows) {
  if (n_q > 0 && (!out_scores || !out_rows)) return kv_fail(KV_ERR_INVALID, "kv_topk: output buffers are NULL");
  return topk_impl(ix, q_indptr, q_ids, q_tf, q_oov_tf2, n_q, k, out_scores, out_rows, nullptr, nullptr);
}

int kv_topk_device(kv_index *ix, const int64_t *q_indptr, const uint32_t *q_ids, const uint32_t *q_tf,
                   const double *q_oov_tf2, int64_t n_q, int k, void *d_scores, void *d_rows) {
  if (n_q > 0 && (!d_scores || !d_rows)) return kv_fail(KV_ERR_INVALID, "kv_topk_device: output buffers are NULL");
  return topk_impl(ix, q_indptr, q_ids, q_tf, q_oov_tf2, n_q, k, nullptr, nullptr, d_scores, d_rows);
}

int kv_merge_topk_device(int device, const void *d_scores_in, const void *d_rows_in, int n_lists, int64_t n_q, int k,
                         void *d_scores_out, void *d_rows_out) {
  if (n_lists < 1 || n_lists > 2048 || n_q < 0 || k < 1 || k > 255 || !d_scores_in || !d_rows_in || !d_scores_out || !d_rows_out)
    return kv_fail(KV_ERR_INVALID, "kv_merge_topk_device: bad arguments");
  if (n_q == 0) return KV_OK;
  KV_CUDA(cudaSetDevice(device));
  merge_topk_kernel<<<(unsigned)((n_q * 32 + 255) / 256), 256>>>((const float *)d_scores_in, (const long long *)d_rows_in,
                                                                 n_lists, n_q, k, nullptr, (float *)d_scores_out,
                                                                 (long long *)d_rows_out);
  KV_CUDA(cudaGetLastError());
  KV_CUDA(cudaDeviceSynchronize());
  return KV_OK;
}

int kv_index_last_timing(const kv_index *ix, float ms[4]) {
  if (!ix || !ms) return kv_fail(KV_ERR_INVALID, "kv_index_last_timing: bad arguments");
  for (int i = 0; i < 4; i++) ms[i] = ix->last_ms[i];
  return KV_OK;
}

int kv_index_last_score_ms(const kv_index *ix, float *ms) {
  if (!ix || !ms) return kv_fail(KV_ERR_INVALID, "kv_index_last_score_ms: bad arguments");
  *ms = ix->last_score_ms;
  return KV_OK;
}

int kv_index_layout(const kv_index *ix, int64_t bytes[4], int64_t counts[17]) {
  if (!ix || !bytes || !counts) return kv_fail(KV_ERR_INVALID, "kv_index_layout: bad arguments");
  bytes[0] = ix->stream_len * 4;
  bytes[1] = ix->n_rows * 4;
  bytes[2] = (ix->n_chunks + 1) * 8;
  bytes[3] = (ix->sum_len + ix->grp_len) * 4 + (ix->n_chunks + 1) * 12;
  counts[0] = ix->stream_len; counts[1] = ix->n_univ; counts[2] = ix->n_rows;
  counts[3] = ix->last_ctas; counts[4] = ix->last_tiles; counts[5] = ix->last_splits;
  counts[6] = ix->batch_h2d_bytes; counts[7] = ix->n_ovf;
  counts[8] = ix->n_chunks; counts[9] = (int64_t)ix->last_stats[0]; counts[10] = (int64_t)ix->last_stats[1];
  counts[11] = (int64_t)ix->last_stats[2];
  counts[12] = (int64_t)ix->last_stats[3];
  for (int i = 0; i < 4; i++) counts[13 + i] = (int64_t)ix->last_stats[4 + i];
  return KV_OK;
}

}  // extern "C"
