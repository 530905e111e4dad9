// Device code of the TF-IDF cosine index: finalize kernels, K1a (float64 full scan), K1b (query
// batch, fused top-k, block-max pruning), K5 (list merge).  Included by tfidf_index.cu only.
//
// Math (restating sklearn text.py:1650-1739 + pairwise.py:1742-1752 as called
// by services/shared/similarity.py:14-20).  The reference refits TF-IDF on [query]+corpus per
// call; with N corpus rows and corpus document frequency df(t):
//   idf_b(t) = ln((N+2)/(df(t)+1)) + 1      feature t of a row that is NOT in the query
//   idf_q(t) = ln((N+2)/(df(t)+2)) + 1      feature t that IS in the query (the fit saw it once more)
//   B_c      = sum_{t in c} (tf_c(t) idf_b(t))^2                        query independent
//   dot      = sum_{t in q∩c} tf_q(t) tf_c(t) a(t),          a(t) = idf_q(t)^2
//   corr     = sum_{t in q∩c} tf_c(t)^2 d(t),                d(t) = idf_q(t)^2 - idf_b(t)^2  (< 0)
//   |q|^2    = sum_{t in q} (tf_q(t) idf_q(t))^2   (out-of-vocabulary features: df = 0)
//   score    = dot / sqrt(|q|^2 (B_c + corr)),  0 when either side has no feature.
//
// Scan layout in HBM (built by finalize; "position" = index of a row in text-sorted order):
//   * rows are sorted by their feature-id sequence (= token order), so rows with similar text are
//     neighbours; perm[position] is the original row;
//   * features present in EVERY local row with one common tf ("universal": the field names of
//     signature_text, fingerprint.py:60-65) are folded into per-query constants;
//   * the remaining entries of all rows form one self-delimiting uint32 stream
//         [31] last entry of its row   [30:5] feature id   [4:0] tf (31 = see overflow table)
//     (a row without entries carries one sentinel entry); chunkptr[] gives the stream offset of
//     every CHUNK_ROWS-th position -- the unit of work distribution and of pruning.  Entries keep
//     the text order of the features, and the longest prefix shared by ALL rows of a chunk is
//     stored once at the head of the chunk ("core", closed by a marker entry): the scan evaluates
//     it once and restarts every row from that state -- each row is still summed in its own
//     entry order, so the result is bit-identical to an unfactored scan;
//   * per chunk a summary pseudo-row (union of the chunk's features with the max tf, in the same
//     entry format) and the smallest row norm: evaluating it like a row yields an upper bound of
//     every score in the chunk (block-max pruning, exact); features that every chunk summary contains
//     with one tf are kept out of the summaries and enter the bounds as per-query constants;
//   * B32/B64: row norms B_c by position.
#pragma once
#include "kv_cuda.cuh"

namespace kvk {

constexpr int CHUNK_ROWS = 64;
constexpr int SUM_GROUP = 16;  // chunk summaries per group in the bound pass (their shared features are evaluated once)
constexpr unsigned long long OVF_GCORE_BASE = 0xD0000000ULL;  // overflow-key position of a summary-group core: BASE + group
constexpr uint32_t FID_BITS = 26;
constexpr uint32_t FID_MASK = (1u << FID_BITS) - 1;
constexpr uint32_t FID_NONE = FID_MASK;  // sentinel feature id (never in a table)
constexpr uint32_t KEY_EMPTY = 0xFFFFFFFFu;
constexpr uint32_t KEY_MULTI = 0x80000000u;  // table key flag: some query of the tile has tf_q > 1
constexpr uint32_t TF_OVF = 31;
constexpr uint32_t FULL = 0xFFFFFFFFu;
constexpr uint32_t PAD_ENTRY = (FID_NONE << 5) | 1u;
constexpr uint32_t FID_CORE = FID_MASK - 1;  // marker entry: end of the chunk's shared prefix ("core")
constexpr uint32_t CORE_ENTRY = (FID_CORE << 5) | 1u;
constexpr unsigned long long OVF_CORE_BASE = 0xC0000000ULL;  // overflow-key position of a chunk core: BASE + chunk
constexpr float PRUNE_SLACK = 1.00002f;   // bound vs threshold comparisons tolerate fp32 rounding
constexpr float FILTER_SLACK = 0.999996f;

// ----------------------------------------------------------------------------------------
// finalize kernels
// ----------------------------------------------------------------------------------------
__global__ void hist_kernel(const uint32_t *__restrict__ ids, const uint16_t *__restrict__ tf, int64_t nnz,
                            uint32_t *cnt, uint32_t *tfmin, uint32_t *tfmax) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x) {
    uint32_t t = ids[p], f = tf[p];
    atomicAdd(&cnt[t], 1u);
    if (tfmin) {
      atomicMin(&tfmin[t], f);
      atomicMax(&tfmax[t], f);
    }
  }
}

struct IdfTables {
  double *a64, *d64, *bb64;
  uint8_t *univ;
  uint32_t *utf;
};

__global__ void idf_kernel(const uint32_t *__restrict__ df, const uint32_t *__restrict__ cnt,
                           const uint32_t *__restrict__ tfmin, const uint32_t *__restrict__ tfmax, int64_t V,
                           int64_t n_total, int64_t n_local, int jaccard, int corpus_fit, IdfTables T) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= V) return;
  // corpus_fit: TF-IDF fitted on the corpus alone (the query is only transformed): one idf for both sides
  double num = (double)(n_total + (corpus_fit ? 1 : 2));
  double ib = jaccard ? 1.0 : log(num / ((double)df[t] + 1.0)) + 1.0;  // Jaccard: every token weighs 1
  double iq = jaccard ? 1.0 : (corpus_fit ? ib : log(num / ((double)df[t] + 2.0)) + 1.0);
  double a = iq * iq, bb = ib * ib;
  T.a64[t] = a; T.d64[t] = a - bb; T.bb64[t] = bb;
  bool u = n_local > 0 && (int64_t)cnt[t] == n_local && tfmin[t] == tfmax[t];
  T.univ[t] = u ? 1 : 0;
  T.utf[t] = u ? tfmin[t] : 0;
}

// one warp per position: B_c and the number of entries the row keeps in the stream
__global__ void rownorm_kernel(const int64_t *__restrict__ indptr, const uint32_t *__restrict__ ids,
                               const uint16_t *__restrict__ tf, const int *__restrict__ perm, int64_t n_rows,
                               const double *__restrict__ bb64, const uint8_t *__restrict__ univ, double *B64,
                               float *B32, int64_t *keep) {
  int64_t pos = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (pos >= n_rows) return;
  const int64_t r = perm ? perm[pos] : pos;
  // B_c is summed in entry order by one lane-strided pass + a fixed shuffle tree: rows with equal
  // text get the same bits
  double b = 0.0;
  int k = 0;
  for (int64_t p = indptr[r] + lane; p < indptr[r + 1]; p += 32) {
    uint32_t t = ids[p];
    double f = (double)tf[p];
    b += f * f * bb64[t];
    k += univ[t] ? 0 : 1;
  }
  for (int o = 16; o; o >>= 1) {
    b += __shfl_xor_sync(FULL, b, o);
    k += __shfl_xor_sync(FULL, k, o);
  }
  if (lane == 0) {
    B64[pos] = b;
    B32[pos] = (float)b;
    if (keep) keep[pos] = k > 0 ? k : 1;
  }
}

__global__ void chunk_meta_kernel(const float *__restrict__ B32, int64_t n_rows, int64_t n_chunks, float *chunk_minB) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= n_chunks) return;
  int64_t r = c * CHUNK_ROWS;
  {
    float m = INFINITY;  // rows without features (B == 0) always score 0: they do not loosen the bound
    for (int64_t i = r; i < r + CHUNK_ROWS && i < n_rows; i++)
      if (B32[i] > 0.f) m = fminf(m, B32[i]);
    chunk_minB[c] = m;
  }
}

// ----------------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash_fid(uint32_t fid, int log_h) { return (fid * 0x9E3779B1u) >> (32 - log_h); }

__device__ uint32_t ovf_lookup(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ vals,
                               int n, int64_t pos, uint32_t fid) {
  unsigned long long key = ((unsigned long long)pos << 32) | fid;
  int lo = 0, hi = n - 1;
  while (lo <= hi) {
    int mid = (lo + hi) >> 1;
    unsigned long long k = keys[mid];
    if (k == key) return vals[mid];
    if (k < key) lo = mid + 1; else hi = mid - 1;
  }
  return TF_OVF;  // unreachable for a consistent index
}

// ----------------------------------------------------------------------------------------
// K1a: one query against every row, float64 (the drop-in SimilarityEngine.score path)
// ----------------------------------------------------------------------------------------
struct ScoreParams {
  const uint32_t *stream;
  const int64_t *chunkptr;
  const int *perm;
  int64_t n_chunks, n_rows;
  const double *B64;
  const unsigned long long *ovf_keys;
  const uint32_t *ovf_vals;
  int n_ovf;
  // query table (global memory): qw[H] = tf_q * a(t), qd[H] = d(t), keys[H]
  const uint32_t *qkeys;
  const double *qw, *qd;
  int log_h;
  int table_in_smem;
  double nq, dotU, corrU;
  int jaccard;
  double *out;  // by ORIGINAL row
};

__global__ void __launch_bounds__(256) tfidf_score_kernel(ScoreParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int H = 1 << P.log_h;
  const uint32_t *keys = P.qkeys;
  const double *qw = P.qw, *qd = P.qd;
  if (P.table_in_smem) {
    double *s_w = (double *)smem_raw;
    double *s_d = s_w + H;
    uint32_t *s_k = (uint32_t *)(s_d + H);
    for (int i = threadIdx.x; i < H; i += blockDim.x) {
      s_k[i] = P.qkeys[i];
      s_w[i] = P.qw[i];
      s_d[i] = P.qd[i];
    }
    __syncthreads();
    keys = s_k; qw = s_w; qd = s_d;
  }
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = warp; c < P.n_chunks; c += n_warps) {
    const int64_t p0 = P.chunkptr[c], p1 = P.chunkptr[c + 1];
    const int64_t pos0 = c * CHUNK_ROWS;
    int row_in = 0;
    // Per-row sums are accumulated strictly in entry order (warp-uniform accumulators), so rows with
    // identical text get bit-identical scores wherever they sit in the stream -- the GFKB handler's
    // stable sort (services/gfkb/app.py:89) then orders duplicate rows exactly like the reference.
    double du = 0.0, dv = 0.0, mine = 0.0;
    double cu = 0.0, cv = 0.0;  // sums over the chunk's shared prefix
    bool in_core = true;        // until the marker (or the first row end) is seen
    for (int64_t p = p0; p < p1; p += 32) {
      const uint32_t e = (p + lane < p1) ? P.stream[p + lane] : PAD_ENTRY;
      const uint32_t fid = (e >> 5) & FID_MASK;
      const uint32_t lastmask = __ballot_sync(FULL, (e >> 31) != 0);
      const uint32_t coremask = __ballot_sync(FULL, fid == FID_CORE);
      const int my_row_off = __popc(lastmask & ((1u << lane) - 1u));
      double wu = 0.0, wv = 0.0;
      bool hit = false;
      if (fid < FID_CORE) {
        uint32_t h = hash_fid(fid, P.log_h);
        for (;;) {
          uint32_t k = keys[h];
          if (k == KEY_EMPTY) break;
          if (k == fid) {
            uint32_t tf = e & 31u;
            if (tf == TF_OVF) {
              // entries before the marker belong to the core (marker and core sit in the first groups of the chunk)
              bool core_entry = in_core && (coremask == 0 || lane < (__ffs(coremask) - 1));  // every chunk has a marker
              tf = ovf_lookup(P.ovf_keys, P.ovf_vals, P.n_ovf,
                              core_entry ? (int64_t)(OVF_CORE_BASE + c) : pos0 + row_in + my_row_off, fid);
            }
            double f = (double)tf;
            wu = f * qw[h];
            wv = f * f * qd[h];
            hit = true;
            break;
          }
          h = (h + 1) & (H - 1);
        }
      }
      const uint32_t hitmask = __ballot_sync(FULL, hit);
      uint32_t ev = hitmask | lastmask | coremask;
      while (ev) {
        const int j = __ffs(ev) - 1;
        ev &= ev - 1;
        if ((hitmask >> j) & 1u) {
          du += __shfl_sync(FULL, wu, j);
          dv += __shfl_sync(FULL, wv, j);
        }
        if ((coremask >> j) & 1u) { cu = du; cv = dv; in_core = false; }
        if ((lastmask >> j) & 1u) {
          in_core = false;
          const double dot = P.dotU + du;
          double sc;
          if (P.jaccard) {
            const double den = P.nq + P.B64[pos0 + row_in] - dot;
            sc = (den > 0.0 && dot != 0.0) ? dot / den : 0.0;
          } else {
            const double den = P.nq * (P.B64[pos0 + row_in] + P.corrU + dv);
            sc = (den > 0.0 && dot != 0.0) ? dot / sqrt(den) : 0.0;
          }
          if ((row_in & 31) == lane) mine = sc;
          row_in++;
          du = cu; dv = cv;
          if ((row_in & 31) == 0) P.out[P.perm[pos0 + row_in - 32 + lane]] = mine;
        }
      }
    }
    if ((row_in & 31) != 0 && lane < (row_in & 31)) P.out[P.perm[pos0 + (row_in & ~31) + lane]] = mine;
  }
}

// ----------------------------------------------------------------------------------------
// K1b: query batch against every row with fused top-k and block-max pruning
// ----------------------------------------------------------------------------------------
struct TileDesc {
  int q_begin, q_count, n_extras, pad;
};

struct TopkParams {
  const uint32_t *stream;
  const int64_t *chunkptr;
  const uint32_t *sum_stream;  // chunk summaries (pseudo-rows)
  const int64_t *sumptr;
  const uint32_t *grp_stream;  // the same summaries in groups of SUM_GROUP: shared core + per-chunk residuals
  const int64_t *grpptr;
  const float *chunk_minB;
  const int *perm;
  int64_t n_chunks, n_rows, row_base;
  const float *B32;
  const unsigned long long *ovf_keys;
  const uint32_t *ovf_vals;
  int n_ovf;
  const unsigned char *tables;  // [n_tiles][table_bytes]
  const TileDesc *tiles;
  const float *q_nq, *q_dotU, *q_corrU;  // [n_q] (sorted query order)
  const float *q_dotS, *q_corrS;         // [n_q] start values of chunk bounds (universal + summary-universal features)
  const int *q_excl;                     // [n_q] or NULL: local ORIGINAL row a query must not match (self-join), -1 = none
  int *gthr;                             // [n_q] float bits: lower bound of the global k-th score
  int *peer_gthr[7];                     // the same array on the other GPUs of a row-sharded GFKB (peer memory over
  int n_peers;                           //   NVLink): a raised bound is pushed to every shard, so all of them prune with it
  int share;                             // thresholds are exchanged (several row splits and/or peers)
  float *ubuf;                           // [n_tiles][n_chunks] chunk upper bounds (scratch)
  unsigned long long *stats;             // [0] chunks scanned, [1] chunks pruned, [2] summaries evaluated
  int64_t n_q;
  int k, n_splits, prune, jaccard;
  float *part_scores;  // [n_splits][n_q][k]
  long long *part_rows;
};

template <int G, int LOGH, int XCAP>
struct TileLayout {
  static constexpr int H = 1 << LOGH;
  static constexpr int QT = 32 * G;
  static constexpr size_t off_keys = 0;
  static constexpr size_t off_ad = off_keys + sizeof(uint32_t) * H;
  static constexpr size_t off_masks = off_ad + sizeof(float2) * H;
  // XCAP (<= 32) extra entries for features some query of the tile holds with tf_q > 1: (weight (t-1) a(t), chain flag)
  // + the membership masks of the queries with that tf_q
  static constexpr size_t off_xad = off_masks + sizeof(uint32_t) * H * G;
  static constexpr size_t off_xmask = off_xad + sizeof(float2) * XCAP;
  static constexpr size_t table_bytes = off_xmask + sizeof(uint32_t) * XCAP * G;  // multiple of 16
  static size_t smem_bytes(int k) { return table_bytes + (size_t)QT * k * 8 + (size_t)QT * 8 + 288; }
};

template <int G>
struct Lanes {  // per-lane state of the G queries a lane owns (query g*32+lane of the tile)
  float nq[G], dotU[G], corrU[G];
  float filt[G], fq[G];  // filter threshold (a score) and the factor of the division-free pre-test
  int krow[G];
  bool valid[G];
  int jaccard;           // 0: TF-IDF cosine, 1: token-set Jaccard (warp-uniform)
};

// score of one (query, row) pair from the accumulated sums.  Cosine: dot / sqrt(|q|^2 (B + corr)).
// Jaccard (a = 1, d = 0, B = |row|, nq = |query|): dot / (|query| + |row| - dot) -- exact small integers.
__device__ __forceinline__ float pair_score(int jaccard, float dot, float nq, float t) {
  if (jaccard) {
    const float den = nq + t - dot;
    return den > 0.f ? __fdiv_rn(dot, den) : 0.f;
  }
  const float den = nq * t;
  return den > 0.f ? __fdiv_rn(dot, __fsqrt_rn(den)) : 0.f;
}
// upper bound from a summary (max dot, min norm): +inf when the denominator bound is not positive
__device__ __forceinline__ float bound_score(int jaccard, float dot, float nq, float t) {
  if (jaccard) {
    const float den = nq + t - dot;
    return den > 0.f ? __fdiv_rn(dot, den) : INFINITY;
  }
  const float den = nq * t;
  return den > 0.f ? __fdiv_rn(dot, __fsqrt_rn(den)) : INFINITY;
}

template <int G>
__device__ __forceinline__ void set_filter(Lanes<G> &L, int g, float ks, int kr) {
  L.filt[g] = ks;
  L.krow[g] = kr;
  // pre-test without division: cosine  dot^2 >= fq * (B + corr);  Jaccard  dot >= fq * (|q| + |row| - dot)
  L.fq[g] = ks > 0.f ? (L.jaccard ? ks * FILTER_SLACK : ks * ks * L.nq[g] * FILTER_SLACK) : -1.f;
}

// shared-memory loads through 32-bit shared-window addresses (the tile table is read-only once it is staged): keeps
// the address arithmetic of the event loop to one multiply-add per load
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float2 lds_f2(uint32_t a) {
  float2 v;
  asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t a) {
  uint4 v;
  asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
// values the compiler must keep in a register instead of re-deriving them inside the event loop
__device__ __forceinline__ uint32_t pinned_lanemask_eq() {
  uint32_t v;
  asm volatile("mov.u32 %0, %%lanemask_eq;" : "=r"(v));
  return v;
}
__device__ __forceinline__ uint32_t pinned_shared_addr(const void *p) {
  uint32_t v;
  asm volatile("mov.u32 %0, %1;" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)));
  return v;
}

// Walk the entries [p0,p1) of one chunk (or of one summary pseudo-row).  For every row end the
// functor gets the per-lane sums (dot, corr include the folded universal features).
//
// Lanes probe 32 stream entries in parallel; the warp then walks the hit / row-end events in stream order (sums
// must follow each row's own entry order).  A feature for which some query of the tile has tf_q > 1 carries the
// KEY_MULTI flag and the index of its first EXTRA entry: the primary slot adds the tf_q = 1 part for every query
// that has the feature, each extra entry adds (t - 1) a(t) tf_c for the queries whose tf_q equals t.
template <int G, int LOGH, int XCAP, bool HAS_CORE, class RowFn>
__device__ __forceinline__ void scan_entries(const uint32_t *__restrict__ stream, int64_t p0, int64_t p1,
                                             int64_t ovf_pos0, int64_t ovf_core_key, uint32_t tbl, uint32_t lanebit,
                                             const unsigned long long *ovf_keys, const uint32_t *ovf_vals, int n_ovf,
                                             const float (&dot0)[G], const float (&corr0)[G], RowFn &&on_row) {
  // tbl: shared-window address of the tile table (TileLayout); lanebit: 1 << lane.
  // HAS_CORE: the range starts with the chunk's shared prefix, closed by a marker entry; the sums
  // reached at the marker are the state every row of the chunk restarts from.
  // Sums are kept for all G groups of the tile -- a group that sits a chunk out is simply not looked at by the row
  // functor (a predicated-off add costs the same issue slot as a skipped one).
  using TL = TileLayout<G, LOGH, XCAP>;
  constexpr int H = 1 << LOGH;
  static_assert(G == 4, "the event loop loads the four membership words of a slot as one uint4");
  const int lane = threadIdx.x & 31;
  float dot[G], corr[G], dotc[G], corrc[G];
#pragma unroll
  for (int g = 0; g < G; g++) { dot[g] = dotc[g] = dot0[g]; corr[g] = corrc[g] = corr0[g]; }
  int row_in = 0;
  bool in_core = HAS_CORE;
  for (int64_t p = p0; p < p1; p += 32) {
    const uint32_t e = (p + lane < p1) ? stream[p + lane] : PAD_ENTRY;
    const uint32_t fid = (e >> 5) & FID_MASK;
    int w = -1;  // (slot << 11) | (first extra << 6) | (multi << 5) | tf  when this lane's entry is in the tile table
    if (fid < FID_CORE) {
      uint32_t h = hash_fid(fid, LOGH);
      for (;;) {
        const uint32_t key = lds_u32(tbl + (uint32_t)TL::off_keys + h * 4u);
        if (key == KEY_EMPTY) break;
        if ((key & FID_MASK) == fid) {
          w = (int)((h << 11) | (((key >> FID_BITS) & 31u) << 6) | ((key >> 31) << 5) | (e & 31u));
          break;
        }
        h = (h + 1) & (H - 1);
      }
    }
    const uint32_t lastmask = __ballot_sync(FULL, (e >> 31) != 0);
    const uint32_t ev_all = __ballot_sync(FULL, w >= 0) | lastmask;
    // the core marker (once per chunk) splits its batch in two: events before it close the shared prefix
    uint32_t ev_first = ev_all, ev_second = 0;
    bool split = false;
    if (HAS_CORE) {
      const uint32_t coremask = __ballot_sync(FULL, fid == FID_CORE);
      if (coremask) {
        const int cj = __ffs(coremask) - 1;
        ev_first = ev_all & ((1u << cj) - 1u);
        ev_second = ev_all & ~((2u << cj) - 1u);
        split = true;
      }
    }
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
      uint32_t ev = pass == 0 ? ev_first : ev_second;
      while (ev) {
        const int j = __ffs(ev) - 1;
        ev &= ev - 1;
        const int wj = __shfl_sync(FULL, w, j);
        if (wj >= 0) {
          const uint32_t slot = (uint32_t)wj >> 11;
          uint32_t tf = wj & 31;
          if (tf == TF_OVF) {
            uint32_t fj = __shfl_sync(FULL, fid, j);
            tf = ovf_lookup(ovf_keys, ovf_vals, n_ovf, in_core ? ovf_core_key : ovf_pos0 + row_in, fj);
          }
          const uint4 mm = lds_u4(tbl + (uint32_t)TL::off_masks + slot * 16u);
          const float2 ad = lds_f2(tbl + (uint32_t)TL::off_ad + slot * 8u);
          const float f = (float)tf;
          const float u = f * ad.x, v = f * f * ad.y;
          if (mm.x & lanebit) { dot[0] += u; corr[0] += v; }
          if (mm.y & lanebit) { dot[1] += u; corr[1] += v; }
          if (mm.z & lanebit) { dot[2] += u; corr[2] += v; }
          if (mm.w & lanebit) { dot[3] += u; corr[3] += v; }
          if (wj & 32) {  // rare: the queries with tf_q = t > 1 get the remaining (t - 1) parts
            uint32_t x = ((uint32_t)wj >> 6) & 31u;
            for (;;) {
              const float2 xa = lds_f2(tbl + (uint32_t)TL::off_xad + x * 8u);
              const uint4 xm = lds_u4(tbl + (uint32_t)TL::off_xmask + x * 16u);
              const float u2 = f * xa.x;
              if (xm.x & lanebit) dot[0] += u2;
              if (xm.y & lanebit) dot[1] += u2;
              if (xm.z & lanebit) dot[2] += u2;
              if (xm.w & lanebit) dot[3] += u2;
              if (xa.y == 0.f) break;  // last extra entry of this feature
              x++;
            }
          }
        }
        if ((lastmask >> j) & 1u) {
          on_row(row_in, dot, corr);
#pragma unroll
          for (int g = 0; g < G; g++) { dot[g] = dotc[g]; corr[g] = corrc[g]; }
          row_in++;
        }
      }
      if (!split) break;
      if (pass == 0) {
#pragma unroll
        for (int g = 0; g < G; g++) { dotc[g] = dot[g]; corrc[g] = corr[g]; }
        in_core = false;
      }
    }
  }
}

template <int G, int LOGH, int XCAP>
__global__ void __launch_bounds__(256, 3) tfidf_topk_kernel(TopkParams P) {
  using TL = TileLayout<G, LOGH, XCAP>;
  constexpr int QT = TL::QT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t tbl = pinned_shared_addr(smem_raw);  // tile table (keys, (a,d), membership masks, extras)
  const uint32_t lanebit = pinned_lanemask_eq();
  float *s_lscore = (float *)(smem_raw + TL::table_bytes);  // [QT][k]
  int *s_lrow = (int *)(s_lscore + QT * P.k);               // [QT][k]
  int *s_cnt = s_lrow + QT * P.k;                           // [QT]
  int *s_lock = s_cnt + QT;                                 // [QT]
  float *s_thrmin = (float *)(s_lock + QT);                 // [1] min over the tile of the k-th scores
  unsigned int *s_stat = (unsigned int *)(s_thrmin + 1);    // [4]
  long long *s_next = (long long *)(s_stat + 5);            // [N_LEVELS + 1] chunk cursors, one per bound level (8-byte aligned: s_thrmin sits on a 16-byte boundary)

  const int tile = blockIdx.x, split = blockIdx.y;
  const TileDesc td = P.tiles[tile];
  const int k = P.k;
  {
    const uint4 *src = (const uint4 *)(P.tables + (size_t)tile * TL::table_bytes);
    uint4 *dst = (uint4 *)smem_raw;
    for (int i = threadIdx.x; i < (int)(TL::table_bytes / 16); i += blockDim.x) dst[i] = src[i];
    for (int i = threadIdx.x; i < QT * k; i += blockDim.x) {
      s_lscore[i] = -INFINITY;
      s_lrow[i] = 0x7fffffff;
    }
    for (int i = threadIdx.x; i < QT; i += blockDim.x) { s_cnt[i] = 0; s_lock[i] = 0; }
    if (threadIdx.x < 4) s_stat[threadIdx.x] = 0;
    if (threadIdx.x == 0) *s_thrmin = -INFINITY;
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  Lanes<G> L;
  L.jaccard = P.jaccard;
#pragma unroll
  for (int g = 0; g < G; g++) {
    int qi = g * 32 + lane;
    L.valid[g] = qi < td.q_count;
    int q = td.q_begin + (L.valid[g] ? qi : 0);
    L.nq[g] = P.q_nq[q];
    L.dotU[g] = P.q_dotU[q];
    L.corrU[g] = P.q_corrU[q];
    if (L.nq[g] <= 0.f) L.valid[g] = false;  // null query (every score is 0): answered on the host side
    set_filter<G>(L, g, L.valid[g] ? __int_as_float(P.gthr[q]) : INFINITY, 0x7fffffff);
  }

  // Publish a lower bound of a query's global k-th score: locally (the CTAs scanning other row ranges) and, when it
  // raises the local value, on every peer GPU (fire-and-forget system-scope reductions over NVLink peer memory).
  // Valid for all shards: k rows with at least this score exist somewhere in the GFKB.
  auto publish_threshold = [&](int64_t q, float ks) {
    const int v = __float_as_int(ks);
    const int old = atomicMax(&P.gthr[q], v);
    if (old < v) {
#pragma unroll
      for (int p = 0; p < 7; p++)
        if (p < P.n_peers) atomicMax_system(P.peer_gthr[p] + q, v);
    }
  };

  // pick up thresholds raised meanwhile by other warps of the CTA (shared lists) and by the CTAs
  // scanning other row ranges for the same queries (global lower bounds of the k-th score)
  auto refresh_filters = [&]() {
#pragma unroll
    for (int g = 0; g < G; g++) {
      int qi = g * 32 + lane;
      if (!L.valid[g]) continue;
      if (s_cnt[qi] == k) {
        float ks = s_lscore[qi * k + k - 1];
        int kr = s_lrow[qi * k + k - 1];
        if (ks > L.filt[g] || (ks == L.filt[g] && kr < L.krow[g])) set_filter<G>(L, g, ks, kr);
      }
      if (P.share) {
        float gs = __int_as_float(*(volatile int *)&P.gthr[td.q_begin + qi]);
        if (gs > L.filt[g]) set_filter<G>(L, g, gs, 0x7fffffff);
      }
    }
  };

  auto process_chunk = [&](int64_t c, uint32_t gmask) {
    const int64_t pos0 = c * CHUNK_ROWS;
    refresh_filters();
    scan_entries<G, LOGH, XCAP, true>(P.stream, P.chunkptr[c], P.chunkptr[c + 1], pos0, (int64_t)(OVF_CORE_BASE + c), tbl, lanebit,
                                P.ovf_keys, P.ovf_vals, P.n_ovf, L.dotU, L.corrU,
                          [&](int row_in, const float *dot, const float *corr) {
      const float Bc = P.B32[pos0 + row_in];
#pragma unroll
      for (int g = 0; g < G; g++) {
        if (!((gmask >> g) & 1u)) continue;
        const float t = Bc + corr[g];
        const float lhs = L.jaccard ? dot[g] : dot[g] * dot[g];
        const float rhs = L.fq[g] * (L.jaccard ? (L.nq[g] + t - dot[g]) : t);
        if (L.valid[g] && lhs >= rhs) {
          const float s = pair_score(L.jaccard, dot[g], L.nq[g], t);
          const int row = P.perm[pos0 + row_in];
          const int qi = g * 32 + lane;
          // self-join: a query never matches the row it was taken from (looked up only on this rare path)
          const bool banned = P.q_excl != nullptr && P.q_excl[td.q_begin + qi] == row;
          if (!banned && (s > L.filt[g] || (s == L.filt[g] && row < L.krow[g]))) {
            while (atomicCAS(&s_lock[qi], 0, 1) != 0) {}
            __threadfence_block();
            float *ls = s_lscore + qi * k;
            int *lr = s_lrow + qi * k;
            int cnt = s_cnt[qi];
            int pos = -1;
            if (cnt < k) {
              pos = cnt;
              s_cnt[qi] = ++cnt;
            } else if (s > ls[k - 1] || (s == ls[k - 1] && row < lr[k - 1])) {
              pos = k - 1;
            }
            if (pos >= 0) {
              while (pos > 0 && (ls[pos - 1] < s || (ls[pos - 1] == s && lr[pos - 1] > row))) {
                ls[pos] = ls[pos - 1];
                lr[pos] = lr[pos - 1];
                pos--;
              }
              ls[pos] = s;
              lr[pos] = row;
            }
            if (cnt == k) {
              float ks = ls[k - 1];
              int kr = lr[k - 1];
              if (ks > L.filt[g] || (ks == L.filt[g] && kr < L.krow[g])) set_filter<G>(L, g, ks, kr);
              if (pos >= 0 && P.share) publish_threshold(td.q_begin + qi, ks);
            }
            __threadfence_block();
            atomicExch(&s_lock[qi], 0);
          }
        }
      }
    });
  };

  long long t_ph1 = 0, t_re = 0, t_scan = 0, t_wait = 0;  // per-warp cycle counters (profiling aid)
  constexpr int N_LEVELS = 24;  // bound levels of the visit order
  // chunks of this split
  const int64_t n_groups = (P.n_chunks + SUM_GROUP - 1) / SUM_GROUP;
  const int64_t g_lo = n_groups * split / P.n_splits, g_hi = n_groups * (split + 1) / P.n_splits;
  const int64_t c_lo = min(P.n_chunks, g_lo * SUM_GROUP), c_hi = min(P.n_chunks, g_hi * SUM_GROUP);

  if (!P.prune) {
    unsigned done = 0;
    for (int64_t c = c_lo + warp; c < c_hi; c += n_warps, done++) process_chunk(c, FULL);
    if (lane == 0) atomicAdd(&s_stat[0], done);
  } else {
    // ---- phase 1: an upper bound of every score in each chunk, from the chunk's summary pseudo-row ----
    float *ub = P.ubuf + (size_t)tile * P.n_chunks;
    long long t0 = clock64();
    for (int64_t gg = g_lo + warp; gg < g_hi; gg += n_warps) {
      float sd[G], sc0[G];  // start of a bound: universal features + the features (nearly) every chunk summary has
#pragma unroll
      for (int g = 0; g < G; g++) {
        const int q = td.q_begin + (L.valid[g] ? g * 32 + lane : 0);
        sd[g] = P.q_dotS[q];
        sc0[g] = P.q_corrS[q];
      }
      // one group = SUM_GROUP chunk summaries: the features they all share are evaluated once (core), then
      // every chunk's residual -> its bound
      scan_entries<G, LOGH, XCAP, true>(P.grp_stream, P.grpptr[gg], P.grpptr[gg + 1], P.n_rows + P.n_chunks + gg * SUM_GROUP,
                                  (int64_t)(OVF_GCORE_BASE + gg), tbl, lanebit,
                                  P.ovf_keys, P.ovf_vals, P.n_ovf, sd, sc0,
                                  [&](int row_in, const float *dot, const float *corr) {
        const int64_t c = gg * SUM_GROUP + row_in;
        const float Bmin = P.chunk_minB[c];
        float best = -INFINITY;
#pragma unroll
        for (int g = 0; g < G; g++) {
          if (!L.valid[g]) continue;
          const float b = bound_score(L.jaccard, dot[g], L.nq[g], Bmin + corr[g]);
          best = fmaxf(best, b);
        }
        for (int o = 16; o; o >>= 1) best = fmaxf(best, __shfl_xor_sync(FULL, best, o));
        if (lane == 0) ub[c] = best;
      });
    }
    t_ph1 = clock64() - t0;
    if (threadIdx.x <= N_LEVELS) s_next[threadIdx.x] = (long long)c_lo;
    t0 = clock64();
    __syncthreads();
    t_wait += clock64() - t0;
    // ---- phase 2: visit chunks by descending bound level; a warp stops once no remaining chunk can beat the
    //      weakest k-th score of the tile.  Every level has its own chunk cursor, so warps move on to the next
    //      level on their own (no barrier, no idle tail per level); each warp derives the tile-wide threshold
    //      from the shared lists itself ----
    auto tile_threshold = [&]() {  // weakest k-th score over the tile's queries (-inf while some query lacks k candidates)
      refresh_filters();
      float m = INFINITY;
#pragma unroll
      for (int g = 0; g < G; g++)
        if (L.valid[g]) m = fminf(m, L.filt[g]);
      for (int o = 16; o; o >>= 1) m = fminf(m, __shfl_xor_sync(FULL, m, o));
      return m;
    };
    for (int lev = 0; lev <= N_LEVELS; lev++) {
      const float hi = lev == 0 ? INFINITY : 1.0f - (lev - 1) * (1.0f / (N_LEVELS - 1));
      const float lo = lev == N_LEVELS ? -INFINITY : 1.0f - lev * (1.0f / (N_LEVELS - 1));
      float thr_min = tile_threshold();
      if (hi * PRUNE_SLACK < thr_min) break;  // every remaining bound is below every k-th score
      for (;;) {
        long long nb = 0;
        if (lane == 0) nb = atomicAdd((unsigned long long *)&s_next[lev], 32ULL);  // warps grab 32 chunks at a time
        const int64_t base = __shfl_sync(FULL, nb, 0);
        if (base >= c_hi) break;
        const int64_t c = base + lane;
        const float b = c < c_hi ? ub[c] : -INFINITY;
        const bool in_level = c < c_hi && b >= lo && (b < hi || lev == 0);  // level 0 takes +inf bounds too
        uint32_t m = __ballot_sync(FULL, in_level && b * PRUNE_SLACK >= thr_min);
        const bool visited = m != 0;
        if (lane == 0) {
          atomicAdd(&s_stat[0], (unsigned)__popc(m));
        }
        while (m) {
          const int j = __ffs(m) - 1;
          m &= m - 1;
          const int64_t cc = base + j;
          // the stored bound is the max over the tile; re-evaluate it per query against each query's own
          // current threshold (much sharper): scan the chunk only if some query could still place a row
          refresh_filters();
          const float Bmin = P.chunk_minB[cc];
          long long t1 = clock64();
          uint32_t may = 0;  // bit g: this lane's query of group g could still place a row of the chunk
          float sd[G], sc0[G];
#pragma unroll
          for (int g = 0; g < G; g++) {
            const int q = td.q_begin + (L.valid[g] ? g * 32 + lane : 0);
            sd[g] = P.q_dotS[q];
            sc0[g] = P.q_corrS[q];
          }
          scan_entries<G, LOGH, XCAP, false>(P.sum_stream, P.sumptr[cc], P.sumptr[cc + 1], P.n_rows + cc, 0, tbl, lanebit,
                                       P.ovf_keys, P.ovf_vals, P.n_ovf, sd, sc0,
                                [&](int, const float *dot, const float *corr) {
#pragma unroll
            for (int g = 0; g < G; g++) {
              if (!L.valid[g]) continue;
              const float b2 = bound_score(L.jaccard, dot[g], L.nq[g], Bmin + corr[g]);
              if (b2 * PRUNE_SLACK >= L.filt[g]) may |= 1u << g;
            }
          });
          uint32_t gmask = 0;
#pragma unroll
          for (int g = 0; g < G; g++)
            if (__any_sync(FULL, (may >> g) & 1u)) gmask |= 1u << g;
          long long t2 = clock64();
          t_re += t2 - t1;
          if (gmask) {
            process_chunk(cc, gmask);
            t_scan += clock64() - t2;
            if (lane == 0) atomicAdd(&s_stat[3], (unsigned)__popc(gmask));
          } else if (lane == 0) {
            atomicAdd(&s_stat[1], 1u);
          }
        }
        if (visited) thr_min = tile_threshold();
      }
    }
    t0 = clock64();
    __syncthreads();
    t_wait += clock64() - t0;
    if (threadIdx.x == 0) s_stat[2] = (unsigned)(c_hi - c_lo);
    if (lane == 0 && P.stats) {
      atomicAdd(&P.stats[4], (unsigned long long)t_ph1);
      atomicAdd(&P.stats[5], (unsigned long long)t_re);
      atomicAdd(&P.stats[6], (unsigned long long)t_scan);
      atomicAdd(&P.stats[7], (unsigned long long)t_wait);
    }
  }
  __syncthreads();
  // publish this CTA's partial lists (already ordered) and raise the global lower bounds
  for (int i = threadIdx.x; i < td.q_count * k; i += blockDim.x) {
    int qi = i / k, j = i - qi * k;
    int64_t q = td.q_begin + qi;
    size_t o = ((size_t)split * P.n_q + q) * k + j;
    bool used = j < s_cnt[qi];
    P.part_scores[o] = used ? s_lscore[i] : -INFINITY;
    P.part_rows[o] = used ? (long long)(P.row_base + s_lrow[i]) : -1LL;
    if (j == k - 1 && used) publish_threshold(q, s_lscore[i]);
  }
  if (threadIdx.x == 0 && P.stats) {
    // s_stat[0]: chunks that passed the tile-wide test, s_stat[1]: of those, rejected per query
    atomicAdd(&P.stats[0], (unsigned long long)(s_stat[0] - s_stat[1]));
    atomicAdd(&P.stats[1], (unsigned long long)((c_hi - c_lo) - (long long)(s_stat[0] - s_stat[1])));
    atomicAdd(&P.stats[2], (unsigned long long)s_stat[2] + s_stat[0]);
    atomicAdd(&P.stats[3], (unsigned long long)s_stat[3]);
  }
}

// ----------------------------------------------------------------------------------------
// K6: float64 re-scoring of selected (query, row) pairs -- one warp per pair.  Walks ALL of the row's raw CSR
// entries in their stored (text) order and sums with explicit round-to-nearest adds/multiplies (no fused
// multiply-add, no folded constants), so the bits depend only on the row's text, the query and the global
// statistics -- not on which segment or shard holds the row or which features that shard folded as universal.
// Rows with identical text therefore tie EXACTLY everywhere, and the (score desc, row asc) order of
// services/gfkb/app.py:89 is reproduced across segments.  Same formula as K1a (values agree to ~1e-16 relative).
// Used by the batched match path: K1b selects candidates in float32, K6 gives them float64 scores.
// ----------------------------------------------------------------------------------------
struct RescoreParams {
  const int64_t *indptr;   // raw CSR of the index (device)
  const uint32_t *ids;
  const uint16_t *tf;
  const double *a64, *d64;  // host-computed idf tables (the values K1a's query tables are built from)
  const double *B64;        // row norms by position
  const int *invperm;       // original row -> position
  const int64_t *q_indptr;  // query batch CSR (device)
  const uint32_t *q_ids, *q_tf;
  const double *q_nq;
  const long long *rows;    // [n_q * k] global row ids (-1: unused slot)
  int64_t n_q, n_rows, row_base, V;
  int k, jaccard;
  double *out;              // [n_q * k]; -inf for unused slots and rows of other shards
};

__global__ void __launch_bounds__(256) rescore_kernel(RescoreParams P) {
  const int64_t pair = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (pair >= P.n_q * P.k) return;
  const int64_t q = pair / P.k;
  const long long gr = P.rows[pair];
  const int64_t r = gr - P.row_base;
  if (gr < 0 || r < 0 || r >= P.n_rows) {
    if (lane == 0) P.out[pair] = -INFINITY;
    return;
  }
  const int64_t qa = P.q_indptr[q], qb = P.q_indptr[q + 1];
  double du = 0.0, dv = 0.0;
  const int64_t p1 = P.indptr[r + 1];
  for (int64_t p0 = P.indptr[r]; p0 < p1; p0 += 32) {
    const int64_t p = p0 + lane;
    bool hit = false;
    double wu = 0.0, wv = 0.0;
    if (p < p1) {
      const uint32_t t = P.ids[p];
      if ((int64_t)t < P.V) {
        for (int64_t j = qa; j < qb; j++)
          if (P.q_ids[j] == t) {
            const double f = (double)P.tf[p];
            wu = __dmul_rn(f, __dmul_rn((double)P.q_tf[j], P.a64[t]));
            wv = __dmul_rn(__dmul_rn(f, f), P.d64[t]);
            hit = true;
            break;
          }
      }
    }
    uint32_t hm = __ballot_sync(FULL, hit);
    while (hm) {
      const int j = __ffs(hm) - 1;
      hm &= hm - 1;
      du = __dadd_rn(du, __shfl_sync(FULL, wu, j));
      dv = __dadd_rn(dv, __shfl_sync(FULL, wv, j));
    }
  }
  if (lane == 0) {
    const double B = P.B64[P.invperm[r]];
    const double dot = du;
    double sc;
    if (P.jaccard) {
      const double den = __dadd_rn(__dadd_rn(P.q_nq[q], B), -dot);
      sc = (den > 0.0 && dot != 0.0) ? dot / den : 0.0;
    } else {
      const double den = __dmul_rn(P.q_nq[q], __dadd_rn(B, dv));
      sc = (den > 0.0 && dot != 0.0) ? dot / sqrt(den) : 0.0;
    }
    P.out[pair] = sc;
  }
}

__global__ void invperm_kernel(const int *__restrict__ perm, int64_t n, int *invperm) {
  const int64_t pos = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (pos < n) invperm[perm[pos]] = (int)pos;
}

// ----------------------------------------------------------------------------------------
// K5: merge n_lists ordered partial lists per query -- one warp per query
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ bool better(float s1, long long r1, float s2, long long r2) {
  // (score desc, row asc); unused slots (-inf, -1) lose against everything real
  if (s1 != s2) return s1 > s2;
  if (r1 < 0) return false;
  if (r2 < 0) return true;
  return r1 < r2;
}

// out_index: optional map from the list's query slot to the output slot (un-sorts the query batch)
__global__ void merge_topk_kernel(const float *__restrict__ in_s, const long long *__restrict__ in_r, int n_lists,
                                  int64_t n_q, int k, const int *__restrict__ out_index, float *out_s,
                                  long long *out_r) {
  const int64_t q = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= n_q) return;
  const int64_t oq = out_index ? out_index[q] : q;
  constexpr int MAXL = 64;  // lists per lane -> up to 2048 lists per launch
  unsigned char head[MAXL];
#pragma unroll
  for (int i = 0; i < MAXL; i++) head[i] = 0;
  for (int j = 0; j < k; j++) {
    float bs = -INFINITY;
    long long br = -1;
    int bl = -1;
    for (int i = 0; i < MAXL; i++) {
      int l = lane + 32 * i;
      if (l >= n_lists) break;
      int h = head[i];
      if (h >= k) continue;
      size_t o = ((size_t)l * n_q + q) * k + h;
      float s = in_s[o];
      long long r = in_r[o];
      if (r >= 0 && (bl < 0 || better(s, r, bs, br))) { bs = s; br = r; bl = l; }
    }
    for (int o = 16; o; o >>= 1) {
      float s2 = __shfl_xor_sync(FULL, bs, o);
      long long r2 = __shfl_xor_sync(FULL, br, o);
      int l2 = __shfl_xor_sync(FULL, bl, o);
      if (l2 >= 0 && (bl < 0 || better(s2, r2, bs, br))) { bs = s2; br = r2; bl = l2; }
    }
    if (bl >= 0 && (bl & 31) == lane) head[bl >> 5]++;
    if (lane == 0) {
      out_s[oq * k + j] = bl >= 0 ? bs : -INFINITY;
      out_r[oq * k + j] = bl >= 0 ? br : -1LL;
    }
  }
}

// Null queries (no feature in common with any row: every score is 0): the stable sort keeps the
// first k rows.  One thread per (query, slot).
__global__ void fill_null_kernel(const int *__restrict__ null_q, int n_null, int k, int64_t n_rows, int64_t row_base,
                                 const int *__restrict__ excl_by_query, float *out_s, long long *out_r) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_null * k) return;
  int q = null_q[i / k], j = i % k;
  const int ex = excl_by_query ? excl_by_query[q] : -1;  // by ORIGINAL query index
  const int64_t r = (ex >= 0 && j >= ex) ? j + 1 : j;     // the j-th row once the excluded one is skipped
  out_s[(size_t)q * k + j] = r < n_rows ? 0.f : -INFINITY;
  out_r[(size_t)q * k + j] = r < n_rows ? row_base + r : -1LL;
}

// Fallback selection for one (irregular) query: k passes of block-wide arg-best over float64 scores.
__global__ void select_topk_kernel(const double *__restrict__ scores, int64_t n, int64_t row_base, int k, int64_t excl,
                                   float *out_s, long long *out_r) {
  __shared__ float s_s[32];
  __shared__ long long s_r[32];
  __shared__ float prev_s;
  __shared__ long long prev_r;
  if (threadIdx.x == 0) { prev_s = INFINITY; prev_r = -1; }
  __syncthreads();
  for (int j = 0; j < k; j++) {
    float bs = -INFINITY;
    long long br = -1;
    const float ps = prev_s;
    const long long pr = prev_r;
    for (int64_t r = threadIdx.x; r < n; r += blockDim.x) {
      if (r == excl) continue;
      float s = (float)scores[r];
      bool after = (s < ps) || (s == ps && r > pr);  // strictly after the previously selected pair
      if (after && (br < 0 || s > bs || (s == bs && r < br))) { bs = s; br = r; }
    }
    for (int o = 16; o; o >>= 1) {
      float s2 = __shfl_xor_sync(FULL, bs, o);
      long long r2 = __shfl_xor_sync(FULL, br, o);
      if (r2 >= 0 && (br < 0 || s2 > bs || (s2 == bs && r2 < br))) { bs = s2; br = r2; }
    }
    if ((threadIdx.x & 31) == 0) { s_s[threadIdx.x >> 5] = bs; s_r[threadIdx.x >> 5] = br; }
    __syncthreads();
    if (threadIdx.x < 32) {
      int nw = blockDim.x >> 5;
      bs = threadIdx.x < nw ? s_s[threadIdx.x] : -INFINITY;
      br = threadIdx.x < nw ? s_r[threadIdx.x] : -1;
      for (int o = 16; o; o >>= 1) {
        float s2 = __shfl_xor_sync(FULL, bs, o);
        long long r2 = __shfl_xor_sync(FULL, br, o);
        if (r2 >= 0 && (br < 0 || s2 > bs || (s2 == bs && r2 < br))) { bs = s2; br = r2; }
      }
      if (threadIdx.x == 0) {
        out_s[j] = br >= 0 ? bs : -INFINITY;
        out_r[j] = br >= 0 ? row_base + br : -1LL;
        if (br >= 0) { prev_s = bs; prev_r = br; } else { prev_s = -INFINITY; prev_r = (long long)n; }
      }
    }
    __syncthreads();
  }
}

}  // namespace kvk
