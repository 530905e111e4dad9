// TF-IDF cosine index for the GFKB match path: finalize (statistics + scan layout),
// K1a (one query -> all float64 scores), K1b (query batch -> fused top-k), K5 (list merge).
//
// Math (restating sklearn text.py:1650-1739 + pairwise.py:1742-1752 as
// called by services/shared/similarity.py:14-20).  The reference refits TF-IDF on
// [query]+corpus per call; with N corpus rows and corpus document frequency df(t):
//   idf_b(t) = ln((N+2)/(df(t)+1)) + 1      feature t of a row that is NOT in the query
//   idf_q(t) = ln((N+2)/(df(t)+2)) + 1      feature t that IS in the query (fit saw it once more)
//   B_c      = sum_{t in c} (tf_c(t) idf_b(t))^2                        query independent
//   dot      = sum_{t in q∩c} tf_q(t) tf_c(t) idf_q(t)^2
//   corr     = sum_{t in q∩c} tf_c(t)^2 (idf_q(t)^2 - idf_b(t)^2)
//   |q|^2    = sum_{t in q} (tf_q(t) idf_q(t))^2   (out-of-vocabulary features: df = 0)
//   score    = dot / sqrt(|q|^2 (B_c + corr)),  0 when either side has no feature.
// a(t) = idf_q(t)^2 and d(t) = idf_q(t)^2 - idf_b(t)^2 depend on t only, so a batch of queries
// never needs a refit: two sums over the q∩c intersection plus a fused epilogue.
//
// Scan layout in HBM (built by finalize):
//   * features present in EVERY local row with one common tf ("universal": the field names
//     of signature_text, fingerprint.py:60-65) are folded into per-query constants and
//     dropped from the rows;
//   * the remaining entries of all rows form one self-delimiting uint32 stream
//         [31] last entry of its row   [30:5] feature id   [4:0] tf (31 = look up in the
//         overflow table; rows without entries carry one sentinel entry)
//     so a warp walks rows without row pointers; chunkptr[] gives the stream offset of every
//     64th row (the unit of work distribution);
//   * B32/B64: row norms B_c.
#include "kv_cuda.cuh"

#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <mutex>
#include <vector>

namespace {

constexpr int CHUNK_ROWS = 64;
constexpr uint32_t FID_BITS = 26;
constexpr uint32_t FID_MASK = (1u << FID_BITS) - 1;
constexpr uint32_t FID_NONE = FID_MASK;  // sentinel feature id (never in a table)
constexpr uint32_t KEY_EMPTY = 0xFFFFFFFFu;
constexpr uint32_t KEY_MULTI = 0x80000000u;  // table key flag: some query of the tile has tf_q > 1
constexpr uint32_t TF_OVF = 31;
constexpr uint32_t FULL = 0xFFFFFFFFu;

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
  float *a32, *d32;
  uint8_t *univ;
  uint32_t *utf;
};

__global__ void idf_kernel(const uint32_t *__restrict__ df, const uint32_t *__restrict__ cnt,
                           const uint32_t *__restrict__ tfmin, const uint32_t *__restrict__ tfmax, int64_t V,
                           int64_t n_total, int64_t n_local, IdfTables T) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= V) return;
  double num = (double)(n_total + 2);
  double ib = log(num / ((double)df[t] + 1.0)) + 1.0;
  double iq = log(num / ((double)df[t] + 2.0)) + 1.0;
  double a = iq * iq, bb = ib * ib, d = a - bb;
  T.a64[t] = a; T.d64[t] = d; T.bb64[t] = bb;
  T.a32[t] = (float)a; T.d32[t] = (float)d;
  bool u = n_local > 0 && (int64_t)cnt[t] == n_local && tfmin[t] == tfmax[t];
  T.univ[t] = u ? 1 : 0;
  T.utf[t] = u ? tfmin[t] : 0;
}

// one warp per row: B_c and the number of entries the row keeps in the stream
__global__ void rownorm_kernel(const int64_t *__restrict__ indptr, const uint32_t *__restrict__ ids,
                               const uint16_t *__restrict__ tf, int64_t n_rows, const double *__restrict__ bb64,
                               const uint8_t *__restrict__ univ, double *B64, float *B32, int64_t *keep) {
  int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= n_rows) return;
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
    B64[r] = b;
    B32[r] = (float)b;
    keep[r] = k > 0 ? k : 1;
  }
}

__global__ void fill_stream_kernel(const int64_t *__restrict__ indptr, const uint32_t *__restrict__ ids,
                                   const uint16_t *__restrict__ tf, int64_t n_rows,
                                   const uint8_t *__restrict__ univ, const int64_t *__restrict__ sptr,
                                   uint32_t *stream, unsigned long long *ovf_keys, uint32_t *ovf_vals,
                                   int *ovf_count, int ovf_cap) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  int64_t w = sptr[r], w_end = sptr[r + 1];
  for (int64_t p = indptr[r]; p < indptr[r + 1]; p++) {
    uint32_t t = ids[p];
    if (univ[t]) continue;
    uint32_t f = tf[p];
    if (f >= TF_OVF) {
      int slot = atomicAdd(ovf_count, 1);
      if (slot < ovf_cap) {
        ovf_keys[slot] = ((unsigned long long)r << 32) | t;
        ovf_vals[slot] = f;
      }
      f = TF_OVF;
    }
    uint32_t last = (w + 1 == w_end) ? 0x80000000u : 0u;
    stream[w++] = last | (t << 5) | f;
  }
  if (w < w_end) stream[w] = 0x80000000u | (FID_NONE << 5) | 1u;  // row without stored entries
}

__global__ void chunkptr_kernel(const int64_t *__restrict__ sptr, int64_t n_rows, int64_t n_chunks,
                                int64_t *chunkptr) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c > n_chunks) return;
  int64_t r = c * CHUNK_ROWS;
  chunkptr[c] = sptr[r < n_rows ? r : n_rows];
}

// ----------------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash_fid(uint32_t fid, int log_h) { return (fid * 0x9E3779B1u) >> (32 - log_h); }

__device__ uint32_t ovf_lookup(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ vals,
                               int n, int64_t row, uint32_t fid) {
  unsigned long long key = ((unsigned long long)row << 32) | fid;
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
  int64_t n_chunks, n_rows;
  const double *B64;
  const unsigned long long *ovf_keys;
  const uint32_t *ovf_vals;
  int n_ovf;
  // query table (global memory): keys[H], wq[H] = tf_q * a(t), dd[H] = d(t)
  const uint32_t *qkeys;
  const double *qw, *qd;
  int log_h;
  int table_in_smem;
  double nq, dotU, corrU;
  double *out;
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
    const int64_t row0 = c * CHUNK_ROWS;
    int row_in = 0;
    // Per-row sums are accumulated strictly in entry order (warp-uniform accumulators), so rows with
    // identical text get bit-identical scores wherever they sit in the stream -- the GFKB handler's
    // stable sort (services/gfkb/app.py:89) then orders duplicate rows exactly like the reference.
    double du = 0.0, dv = 0.0, mine = 0.0;
    for (int64_t p = p0; p < p1; p += 32) {
      const uint32_t e = (p + lane < p1) ? P.stream[p + lane] : ((FID_NONE << 5) | 1u);
      const uint32_t fid = (e >> 5) & FID_MASK;
      const uint32_t lastmask = __ballot_sync(FULL, (e >> 31) != 0);
      const int my_row_off = __popc(lastmask & ((1u << lane) - 1u));
      double wu = 0.0, wv = 0.0;
      bool hit = false;
      if (fid != FID_NONE) {
        uint32_t h = hash_fid(fid, P.log_h);
        for (;;) {
          uint32_t k = keys[h];
          if (k == KEY_EMPTY) break;
          if (k == fid) {
            uint32_t tf = e & 31u;
            if (tf == TF_OVF) tf = ovf_lookup(P.ovf_keys, P.ovf_vals, P.n_ovf, row0 + row_in + my_row_off, fid);
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
      uint32_t ev = hitmask | lastmask;
      while (ev) {
        const int j = __ffs(ev) - 1;
        ev &= ev - 1;
        if ((hitmask >> j) & 1u) {
          du += __shfl_sync(FULL, wu, j);
          dv += __shfl_sync(FULL, wv, j);
        }
        if ((lastmask >> j) & 1u) {
          const int64_t r = row0 + row_in;
          const double dot = P.dotU + du;
          const double den = P.nq * (P.B64[r] + P.corrU + dv);
          const double sc = (den > 0.0 && dot != 0.0) ? dot / sqrt(den) : 0.0;
          if ((row_in & 31) == lane) mine = sc;
          row_in++;
          du = 0.0; dv = 0.0;
          if ((row_in & 31) == 0) P.out[row0 + row_in - 32 + lane] = mine;  // coalesced store of 32 rows
        }
      }
    }
    if ((row_in & 31) != 0 && lane < (row_in & 31)) P.out[row0 + (row_in & ~31) + lane] = mine;
  }
}

// ----------------------------------------------------------------------------------------
// K1b: query batch against every row with fused top-k
// ----------------------------------------------------------------------------------------
struct TileDesc {
  int q_begin, q_count, n_extras, pad;
};

struct TopkParams {
  const uint32_t *stream;
  const int64_t *chunkptr;
  int64_t n_chunks, n_rows, row_base;
  const float *B32;
  const unsigned long long *ovf_keys;
  const uint32_t *ovf_vals;
  int n_ovf;
  const unsigned char *tables;  // [n_tiles][table_bytes]
  const TileDesc *tiles;
  const float *q_nq, *q_dotU, *q_corrU;  // [n_q]
  int *gthr;                             // [n_q] float bits: lower bound of the global k-th score
  int64_t n_q;
  int k, n_splits;
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
  static constexpr size_t off_xkey = off_masks + sizeof(uint32_t) * H * G;
  static constexpr size_t off_xtf = off_xkey + sizeof(uint32_t) * XCAP;
  static constexpr size_t table_bytes = off_xtf + sizeof(float) * XCAP;  // multiple of 16
  static size_t smem_bytes(int k) { return table_bytes + (size_t)QT * k * 8 + (size_t)QT * 8; }
};

template <int G, int LOGH, int XCAP>
__global__ void __launch_bounds__(256) tfidf_topk_kernel(TopkParams P) {
  using L = TileLayout<G, LOGH, XCAP>;
  constexpr int H = L::H, QT = L::QT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t *s_keys = (uint32_t *)(smem_raw + L::off_keys);
  float2 *s_ad = (float2 *)(smem_raw + L::off_ad);
  uint32_t *s_masks = (uint32_t *)(smem_raw + L::off_masks);
  uint32_t *s_xkey = (uint32_t *)(smem_raw + L::off_xkey);
  float *s_xtf = (float *)(smem_raw + L::off_xtf);
  float *s_lscore = (float *)(smem_raw + L::table_bytes);  // [QT][k]
  int *s_lrow = (int *)(s_lscore + QT * P.k);               // [QT][k]
  int *s_cnt = s_lrow + QT * P.k;                           // [QT]
  int *s_lock = s_cnt + QT;                                 // [QT]

  const int tile = blockIdx.x, split = blockIdx.y;
  const TileDesc td = P.tiles[tile];
  const int k = P.k;
  {
    const uint4 *src = (const uint4 *)(P.tables + (size_t)tile * L::table_bytes);
    uint4 *dst = (uint4 *)smem_raw;
    for (int i = threadIdx.x; i < (int)(L::table_bytes / 16); i += blockDim.x) dst[i] = src[i];
    for (int i = threadIdx.x; i < QT * k; i += blockDim.x) {
      s_lscore[i] = -INFINITY;
      s_lrow[i] = 0x7fffffff;
    }
    for (int i = threadIdx.x; i < QT; i += blockDim.x) { s_cnt[i] = 0; s_lock[i] = 0; }
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  float nq[G], dotU[G], corrU[G];
  float filt[G], fq[G];  // filter threshold (a score) and its squared-domain factor
  int krow[G];
  bool valid[G];
#pragma unroll
  for (int g = 0; g < G; g++) {
    int qi = g * 32 + lane;
    valid[g] = qi < td.q_count;
    int q = td.q_begin + (valid[g] ? qi : 0);
    nq[g] = P.q_nq[q];
    dotU[g] = P.q_dotU[q];
    corrU[g] = P.q_corrU[q];
    filt[g] = valid[g] ? __int_as_float(P.gthr[q]) : INFINITY;
    krow[g] = 0x7fffffff;
    fq[g] = filt[g] > 0.f ? filt[g] * filt[g] * nq[g] * 0.999996f : -1.f;
  }

  // rows of this split, in chunks of 64 rows, round-robin over the CTA's warps
  const int64_t c_lo = P.n_chunks * split / P.n_splits, c_hi = P.n_chunks * (split + 1) / P.n_splits;
  for (int64_t c = c_lo + warp; c < c_hi; c += n_warps) {
    const int64_t p0 = P.chunkptr[c], p1 = P.chunkptr[c + 1];
    const int64_t row0 = c * CHUNK_ROWS;
    int row_in = 0;
    const float Bv0 = (row0 + lane < P.n_rows) ? P.B32[row0 + lane] : 0.f;
    const float Bv1 = (row0 + 32 + lane < P.n_rows) ? P.B32[row0 + 32 + lane] : 0.f;
    // pick up thresholds other warps of the CTA have raised meanwhile
#pragma unroll
    for (int g = 0; g < G; g++) {
      int qi = g * 32 + lane;
      if (valid[g] && s_cnt[qi] == k) {
        float ks = s_lscore[qi * k + k - 1];
        int kr = s_lrow[qi * k + k - 1];
        if (ks > filt[g] || (ks == filt[g] && kr < krow[g])) {
          filt[g] = ks; krow[g] = kr;
          fq[g] = ks > 0.f ? ks * ks * nq[g] * 0.999996f : -1.f;
        }
      }
    }
    float dot[G], corr[G];
#pragma unroll
    for (int g = 0; g < G; g++) { dot[g] = dotU[g]; corr[g] = corrU[g]; }

    for (int64_t p = p0; p < p1; p += 32) {
      const uint32_t e = (p + lane < p1) ? P.stream[p + lane] : ((FID_NONE << 5) | 1u);
      const uint32_t fid = (e >> 5) & FID_MASK;
      int w = -1;  // (slot << 6) | (multi << 5) | tf  when this lane's entry is in the tile table
      if (fid != FID_NONE) {
        uint32_t h = hash_fid(fid, LOGH);
        for (;;) {
          uint32_t key = s_keys[h];
          if (key == KEY_EMPTY) break;
          if ((key & FID_MASK) == fid) {
            w = (int)((h << 6) | ((key >> 31) << 5) | (e & 31u));
            break;
          }
          h = (h + 1) & (H - 1);
        }
      }
      const uint32_t lastmask = __ballot_sync(FULL, (e >> 31) != 0);
      uint32_t ev = __ballot_sync(FULL, w >= 0) | lastmask;
      while (ev) {
        const int j = __ffs(ev) - 1;
        ev &= ev - 1;
        const int wj = __shfl_sync(FULL, w, j);
        if (wj >= 0) {
          const int slot = wj >> 6;
          uint32_t tf = wj & 31;
          if (tf == TF_OVF) {
            uint32_t fj = __shfl_sync(FULL, fid, j);
            tf = ovf_lookup(P.ovf_keys, P.ovf_vals, P.n_ovf, row0 + row_in, fj);
          }
          const float2 ad = s_ad[slot];
          const float f = (float)tf;
          const float u = f * ad.x, v = f * f * ad.y;
          uint32_t m[G];
          if (G == 4) {
            uint4 mm = *(const uint4 *)(s_masks + slot * 4);
            m[0] = mm.x; m[1 % G] = mm.y; m[2 % G] = mm.z; m[3 % G] = mm.w;
          } else {
#pragma unroll
            for (int g = 0; g < G; g++) m[g] = s_masks[slot * G + g];
          }
          if (!(wj & 32)) {
#pragma unroll
            for (int g = 0; g < G; g++)
              if ((m[g] >> lane) & 1u) { dot[g] += u; corr[g] += v; }
          } else {
            // some query of this tile has tf_q > 1 for this feature: fetch per-query multipliers
            float mul[G];
#pragma unroll
            for (int g = 0; g < G; g++) mul[g] = 1.f;
            for (int x = 0; x < td.n_extras; x++) {
              uint32_t xk = s_xkey[x];
              if ((int)(xk >> 8) == slot) {
                int qi = xk & 255;
#pragma unroll
                for (int g = 0; g < G; g++)
                  if (qi == g * 32 + lane) mul[g] = s_xtf[x];
              }
            }
#pragma unroll
            for (int g = 0; g < G; g++)
              if ((m[g] >> lane) & 1u) { dot[g] += mul[g] * u; corr[g] += v; }
          }
        }
        if ((lastmask >> j) & 1u) {
          // ---- fused epilogue for row (row0 + row_in) ----
          const float bsel = (row_in & 32) ? Bv1 : Bv0;
          const float Bc = __shfl_sync(FULL, bsel, row_in & 31);
          const int row = (int)(row0 + row_in);
#pragma unroll
          for (int g = 0; g < G; g++) {
            const float t = Bc + corr[g];
            const float lhs = dot[g] * dot[g];
            const bool pass = valid[g] && (lhs >= fq[g] * t);
            if (pass) {
              const float den = nq[g] * t;
              const float s = (den > 0.f) ? __fdiv_rn(dot[g], __fsqrt_rn(den)) : 0.f;
              if (s > filt[g] || (s == filt[g] && row < krow[g])) {
                const int qi = g * 32 + lane;
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
                  if (ks > filt[g] || (ks == filt[g] && kr < krow[g])) {
                    filt[g] = ks; krow[g] = kr;
                    fq[g] = ks > 0.f ? ks * ks * nq[g] * 0.999996f : -1.f;
                  }
                }
                __threadfence_block();
                atomicExch(&s_lock[qi], 0);
              }
            }
            dot[g] = dotU[g];
            corr[g] = corrU[g];
          }
          row_in++;
        }
      }
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
    if (j == k - 1 && used) atomicMax(&P.gthr[q], __float_as_int(s_lscore[i]));
  }
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

__global__ void merge_topk_kernel(const float *__restrict__ in_s, const long long *__restrict__ in_r, int n_lists,
                                  int64_t n_q, int k, float *out_s, long long *out_r) {
  const int64_t q = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= n_q) return;
  // lane owns lists lane, lane+32, ...; head[] positions kept in a bounded register array
  constexpr int MAXL = 64;  // lists per lane -> up to 2048 lists per launch
  unsigned char head[MAXL];
#pragma unroll
  for (int i = 0; i < MAXL; i++) head[i] = 0;
  for (int j = 0; j < k; j++) {
    float bs = -INFINITY;
    long long br = -1;
    int bl = -1;
#pragma unroll 4
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
    // warp arg-best
    for (int o = 16; o; o >>= 1) {
      float s2 = __shfl_xor_sync(FULL, bs, o);
      long long r2 = __shfl_xor_sync(FULL, br, o);
      int l2 = __shfl_xor_sync(FULL, bl, o);
      if (l2 >= 0 && (bl < 0 || better(s2, r2, bs, br))) { bs = s2; br = r2; bl = l2; }
    }
    if (bl >= 0 && (bl & 31) == lane) {
      int i = bl >> 5;
#pragma unroll
      for (int x = 0; x < MAXL; x++)
        if (x == i) head[x]++;
    }
    if (lane == 0) {
      out_s[q * k + j] = bl >= 0 ? bs : -INFINITY;
      out_r[q * k + j] = bl >= 0 ? br : -1LL;
    }
  }
}

// Fallback selection for one (irregular) query: k passes of block-wide arg-best over float64 scores.
__global__ void select_topk_kernel(const double *__restrict__ scores, int64_t n, int64_t row_base, int k,
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

// tile shape used by the batched scan
constexpr int TG = 4, TLOGH = 11, TXCAP = 256;
using Tile = TileLayout<TG, TLOGH, TXCAP>;

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

  // raw CSR (device)
  DevVec<int64_t> indptr;  // n_rows + 1 entries once any row exists
  DevVec<uint32_t> ids;
  DevVec<uint16_t> tf;
  int64_t n_rows = 0, nnz = 0;

  bool has_gdf = false;
  std::vector<uint32_t> h_gdf;
  int64_t n_global = 0;

  bool finalized = false;
  int64_t V = 0, n_total = 0;
  DevBuf<uint32_t> d_df, d_cnt, d_tfmin, d_tfmax, d_utf;
  DevBuf<double> d_a64, d_d64, d_bb64, d_B64;
  DevBuf<float> d_a32, d_d32, d_B32;
  DevBuf<uint8_t> d_univ;
  DevBuf<int64_t> d_keep, d_sptr, d_chunkptr;
  DevBuf<uint32_t> d_stream;
  DevBuf<unsigned long long> d_ovf_keys, d_ovf_keys2;
  DevBuf<uint32_t> d_ovf_vals, d_ovf_vals2;
  DevBuf<unsigned char> d_cub;
  int n_ovf = 0;
  int64_t stream_len = 0, n_chunks = 0;
  std::vector<uint32_t> h_df;
  std::vector<uint8_t> h_univ;
  std::vector<uint32_t> h_utf;
  int64_t n_univ = 0;

  // query scratch
  PinnedBuf<unsigned char> h_tables;
  PinnedBuf<TileDesc> h_tiles;
  PinnedBuf<float> h_qconst;  // 3 * n_q
  DevBuf<unsigned char> d_tables;
  DevBuf<TileDesc> d_tiles;
  DevBuf<float> d_qconst;
  DevBuf<int> d_gthr;
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
  int64_t batch_q = 0, batch_tiles = 0, batch_h2d_bytes = 0;
  std::vector<int64_t> irr_q, irr_indptr;
  std::vector<uint32_t> irr_ids, irr_tf;
  std::vector<double> irr_oov;

  float last_ms[4] = {0, 0, 0, 0};
  int64_t last_ctas = 0, last_tiles = 0, last_splits = 0;
};

namespace {

struct QueryPrep {
  double nq = 0, dotU = 0, corrU = 0;
  std::vector<uint32_t> fid;  // non-universal, in-vocabulary features
  std::vector<uint32_t> tfq;
};

inline void idf_host(int64_t n_total, uint32_t df, double &a, double &d) {
  double num = (double)(n_total + 2);
  double ib = std::log(num / ((double)df + 1.0)) + 1.0;
  double iq = std::log(num / ((double)df + 2.0)) + 1.0;
  a = iq * iq;
  d = a - ib * ib;
}

int prep_query(const kv_index *ix, const uint32_t *ids, const uint32_t *tf, int64_t nnz, double oov_tf2,
               QueryPrep &out) {
  out.nq = out.dotU = out.corrU = 0;
  out.fid.clear();
  out.tfq.clear();
  double idf0 = std::log((double)(ix->n_total + 2) / 2.0) + 1.0;  // df == 0 features of the query
  out.nq = oov_tf2 * idf0 * idf0;
  for (int64_t i = 0; i < nnz; i++) {
    uint32_t t = ids[i];
    double f = (double)tf[i];
    if ((int64_t)t >= ix->V) {  // id issued after finalize: not in any local row, df from global table or 0
      out.nq += f * f * idf0 * idf0;
      continue;
    }
    double a, d;
    idf_host(ix->n_total, ix->h_df[t], a, d);
    out.nq += f * f * a;
    if (ix->h_univ[t]) {
      double u = (double)ix->h_utf[t];
      out.dotU += f * u * a;
      out.corrU += u * u * d;
    } else {
      out.fid.push_back(t);
      out.tfq.push_back(tf[i]);
    }
  }
  return KV_OK;
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
  ix->d_a32.release(); ix->d_d32.release(); ix->d_B32.release(); ix->d_univ.release();
  ix->d_keep.release(); ix->d_sptr.release(); ix->d_chunkptr.release(); ix->d_stream.release();
  ix->d_ovf_keys.release(); ix->d_ovf_keys2.release(); ix->d_ovf_vals.release(); ix->d_ovf_vals2.release();
  ix->d_cub.release();
  ix->h_tables.release(); ix->h_tiles.release(); ix->h_qconst.release();
  ix->d_tables.release(); ix->d_tiles.release(); ix->d_qconst.release(); ix->d_gthr.release();
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
  std::vector<int64_t> ip((size_t)n_rows + 1);
  for (int64_t i = 0; i <= n_rows; i++) {
    ip[(size_t)i] = indptr[i] - indptr[0] + ix->nnz;
    if (i && ip[(size_t)i] < ip[(size_t)i - 1]) return kv_fail(KV_ERR_INVALID, "kv_index_append: indptr not monotone");
  }
  std::vector<uint16_t> tf16((size_t)add);
  const uint32_t *tfs = tf + indptr[0];
  for (int64_t i = 0; i < add; i++) {
    if (tfs[i] == 0 || tfs[i] > 65535u)
      return kv_fail(KV_ERR_INVALID, "kv_index_append: term frequency %u outside 1..65535", tfs[i]);
    tf16[(size_t)i] = (uint16_t)tfs[i];
  }
  KV_CUDA(ix->indptr.reserve(ix->n_rows + n_rows + 1, ix->stream));
  KV_CUDA(ix->ids.reserve(ix->nnz + add, ix->stream));
  KV_CUDA(ix->tf.reserve(ix->nnz + add, ix->stream));
  // indptr[0..n_rows_old] already there (entry n_rows_old == nnz_old == ip[0])
  KV_CUDA(cudaMemcpyAsync(ix->indptr.p + ix->n_rows, ip.data(), (size_t)(n_rows + 1) * sizeof(int64_t),
                          cudaMemcpyHostToDevice, ix->stream));
  if (add) {
    KV_CUDA(cudaMemcpyAsync(ix->ids.p + ix->nnz, ids + indptr[0], (size_t)add * sizeof(uint32_t),
                            cudaMemcpyHostToDevice, ix->stream));
    KV_CUDA(cudaMemcpyAsync(ix->tf.p + ix->nnz, tf16.data(), (size_t)add * sizeof(uint16_t),
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
  KV_CUDA(ix->d_cnt.ensure(vocab_size));
  KV_CUDA(cudaMemsetAsync(ix->d_cnt.p, 0, (size_t)vocab_size * sizeof(uint32_t), ix->stream));
  if (ix->nnz) {
    // ids must be < vocab_size; verified by finalize, here the caller vouches for its own vocabulary
    hist_kernel<<<ix->sm_count * 8, 256, 0, ix->stream>>>(ix->ids.p, ix->tf.p, ix->nnz, ix->d_cnt.p, nullptr, nullptr);
    KV_CUDA(cudaGetLastError());
  }
  KV_CUDA(cudaMemcpyAsync(df_out, ix->d_cnt.p, (size_t)vocab_size * sizeof(uint32_t), cudaMemcpyDeviceToHost, ix->stream));
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
  if (ix->nnz) {  // every id must be inside the vocabulary
    uint32_t *d_max = nullptr;
    size_t tmp = 0;
    KV_CUDA(cudaMalloc(&d_max, sizeof(uint32_t)));
    cub::DeviceReduce::Max(nullptr, tmp, ix->ids.p, d_max, ix->nnz, s);
    KV_CUDA(ix->d_cub.ensure((int64_t)tmp));
    cub::DeviceReduce::Max(ix->d_cub.p, tmp, ix->ids.p, d_max, ix->nnz, s);
    uint32_t h_max = 0;
    KV_CUDA(cudaMemcpyAsync(&h_max, d_max, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    KV_CUDA(cudaStreamSynchronize(s));
    cudaFree(d_max);
    if ((int64_t)h_max >= V)
      return kv_fail(KV_ERR_INVALID, "kv_index_finalize: feature id %u outside vocabulary of %lld", h_max, (long long)V);
  }
  const int64_t Vz = V > 0 ? V : 1;
  KV_CUDA(ix->d_df.ensure(Vz)); KV_CUDA(ix->d_cnt.ensure(Vz)); KV_CUDA(ix->d_tfmin.ensure(Vz));
  KV_CUDA(ix->d_tfmax.ensure(Vz)); KV_CUDA(ix->d_utf.ensure(Vz));
  KV_CUDA(ix->d_a64.ensure(Vz)); KV_CUDA(ix->d_d64.ensure(Vz)); KV_CUDA(ix->d_bb64.ensure(Vz));
  KV_CUDA(ix->d_a32.ensure(Vz)); KV_CUDA(ix->d_d32.ensure(Vz)); KV_CUDA(ix->d_univ.ensure(Vz));
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
    IdfTables T{ix->d_a64.p, ix->d_d64.p, ix->d_bb64.p, ix->d_a32.p, ix->d_d32.p, ix->d_univ.p, ix->d_utf.p};
    idf_kernel<<<(unsigned)((V + 255) / 256), 256, 0, s>>>(ix->d_df.p, ix->d_cnt.p, ix->d_tfmin.p, ix->d_tfmax.p, V,
                                                            ix->n_total, n, T);
    KV_CUDA(cudaGetLastError());
  }
  const int64_t nz = n > 0 ? n : 1;
  KV_CUDA(ix->d_B64.ensure(nz)); KV_CUDA(ix->d_B32.ensure(nz));
  KV_CUDA(ix->d_keep.ensure(nz + 1)); KV_CUDA(ix->d_sptr.ensure(nz + 1));
  ix->n_chunks = (n + CHUNK_ROWS - 1) / CHUNK_ROWS;
  KV_CUDA(ix->d_chunkptr.ensure(ix->n_chunks + 1));
  ix->stream_len = 0;
  ix->n_ovf = 0;
  if (n) {
    rownorm_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, s>>>(ix->indptr.p, ix->ids.p, ix->tf.p, n, ix->d_bb64.p,
                                                                    ix->d_univ.p, ix->d_B64.p, ix->d_B32.p, ix->d_keep.p);
    KV_CUDA(cudaGetLastError());
    KV_CUDA(cudaMemsetAsync(ix->d_keep.p + n, 0, sizeof(int64_t), s));
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, ix->d_keep.p, ix->d_sptr.p, n + 1, s);
    KV_CUDA(ix->d_cub.ensure((int64_t)tmp));
    cub::DeviceScan::ExclusiveSum(ix->d_cub.p, tmp, ix->d_keep.p, ix->d_sptr.p, n + 1, s);
    KV_CUDA(cudaMemcpyAsync(&ix->stream_len, ix->d_sptr.p + n, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    KV_CUDA(cudaStreamSynchronize(s));
    KV_CUDA(ix->d_stream.ensure(ix->stream_len + 32));
    int ovf_cap = 1 << 16;
    for (;;) {
      KV_CUDA(ix->d_ovf_keys.ensure(ovf_cap)); KV_CUDA(ix->d_ovf_vals.ensure(ovf_cap));
      int *d_count = nullptr;
      KV_CUDA(cudaMalloc(&d_count, sizeof(int)));
      KV_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), s));
      fill_stream_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ix->indptr.p, ix->ids.p, ix->tf.p, n, ix->d_univ.p,
                                                                      ix->d_sptr.p, ix->d_stream.p, ix->d_ovf_keys.p,
                                                                      ix->d_ovf_vals.p, d_count, ovf_cap);
      KV_CUDA(cudaGetLastError());
      int h_count = 0;
      KV_CUDA(cudaMemcpyAsync(&h_count, d_count, sizeof(int), cudaMemcpyDeviceToHost, s));
      KV_CUDA(cudaStreamSynchronize(s));
      cudaFree(d_count);
      if (h_count <= ovf_cap) { ix->n_ovf = h_count; break; }
      ovf_cap = h_count;
    }
    if (ix->n_ovf > 1) {
      KV_CUDA(ix->d_ovf_keys2.ensure(ix->n_ovf)); KV_CUDA(ix->d_ovf_vals2.ensure(ix->n_ovf));
      size_t t2 = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, t2, ix->d_ovf_keys.p, ix->d_ovf_keys2.p, ix->d_ovf_vals.p,
                                      ix->d_ovf_vals2.p, ix->n_ovf, 0, 64, s);
      KV_CUDA(ix->d_cub.ensure((int64_t)t2));
      cub::DeviceRadixSort::SortPairs(ix->d_cub.p, t2, ix->d_ovf_keys.p, ix->d_ovf_keys2.p, ix->d_ovf_vals.p,
                                      ix->d_ovf_vals2.p, ix->n_ovf, 0, 64, s);
      std::swap(ix->d_ovf_keys.p, ix->d_ovf_keys2.p); std::swap(ix->d_ovf_keys.cap, ix->d_ovf_keys2.cap);
      std::swap(ix->d_ovf_vals.p, ix->d_ovf_vals2.p); std::swap(ix->d_ovf_vals.cap, ix->d_ovf_vals2.cap);
    }
    chunkptr_kernel<<<(unsigned)((ix->n_chunks + 1 + 255) / 256), 256, 0, s>>>(ix->d_sptr.p, n, ix->n_chunks, ix->d_chunkptr.p);
    KV_CUDA(cudaGetLastError());
  } else {
    KV_CUDA(ix->d_stream.ensure(32));
    KV_CUDA(cudaMemsetAsync(ix->d_chunkptr.p, 0, sizeof(int64_t), s));
  }
  ix->h_df.assign((size_t)V, 0);
  ix->h_univ.assign((size_t)V, 0);
  ix->h_utf.assign((size_t)V, 0);
  if (V) {
    KV_CUDA(cudaMemcpyAsync(ix->h_df.data(), ix->d_df.p, (size_t)V * 4, cudaMemcpyDeviceToHost, s));
    KV_CUDA(cudaMemcpyAsync(ix->h_univ.data(), ix->d_univ.p, (size_t)V, cudaMemcpyDeviceToHost, s));
    KV_CUDA(cudaMemcpyAsync(ix->h_utf.data(), ix->d_utf.p, (size_t)V * 4, cudaMemcpyDeviceToHost, s));
  }
  KV_CUDA(cudaStreamSynchronize(s));
  ix->n_univ = 0;
  for (uint8_t u : ix->h_univ) ix->n_univ += u;
  ix->V = V;
  ix->finalized = true;
  ix->batch_valid = false;
  return KV_OK;
}

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
    idf_host(ix->n_total, ix->h_df[t], a, d);
    tk[h] = t;
    tw[h] = (double)qp.tfq[i] * a;
    tdd[h] = d;
  }
  cudaStream_t s = ix->stream;
  KV_CUDA(cudaMemcpyAsync(ix->d_qtab.p, ix->h_qtab.p, tab_bytes, cudaMemcpyHostToDevice, s));
  KV_CUDA(ix->d_scores.ensure(ix->n_rows));
  ScoreParams P;
  P.stream = ix->d_stream.p; P.chunkptr = ix->d_chunkptr.p; P.n_chunks = ix->n_chunks; P.n_rows = ix->n_rows;
  P.B64 = ix->d_B64.p; P.ovf_keys = ix->d_ovf_keys.p; P.ovf_vals = ix->d_ovf_vals.p; P.n_ovf = ix->n_ovf;
  P.qw = (const double *)ix->d_qtab.p; P.qd = P.qw + H; P.qkeys = (const uint32_t *)(P.qd + H);
  P.log_h = log_h; P.table_in_smem = tab_bytes <= 40 * 1024;
  P.nq = qp.nq; P.dotU = qp.dotU; P.corrU = qp.corrU; P.out = ix->d_scores.p;
  int64_t warps_needed = ix->n_chunks;
  int blocks = (int)std::min<int64_t>((warps_needed + 7) / 8, (int64_t)ix->sm_count * 8);
  if (blocks < 1) blocks = 1;
  tfidf_score_kernel<<<blocks, 256, P.table_in_smem ? tab_bytes : 0, s>>>(P);
  KV_CUDA(cudaGetLastError());
  if (out_scores)
    KV_CUDA(cudaMemcpyAsync(out_scores, ix->d_scores.p, (size_t)ix->n_rows * sizeof(double), cudaMemcpyDeviceToHost, s));
  KV_CUDA(cudaStreamSynchronize(s));
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
  const int QT = Tile::QT, H = Tile::H;
  ix->batch_valid = false;
  ix->irr_q.clear(); ix->irr_indptr.assign(1, 0); ix->irr_ids.clear(); ix->irr_tf.clear(); ix->irr_oov.clear();

  // host: per-query constants, tile packing, tile tables
  KV_CUDA(ix->h_qconst.ensure(4 * n_q));
  float *c_nq = ix->h_qconst.p, *c_dotU = c_nq + n_q, *c_corrU = c_dotU + n_q, *c_ninf = c_corrU + n_q;
  std::vector<QueryPrep> qp((size_t)n_q);
  std::vector<char> is_irr((size_t)n_q, 0);
  std::vector<TileDesc> tiles;
  {
    TileDesc cur{0, 0, 0, 0};
    int cur_feats = 0;
    for (int64_t q = 0; q < n_q; q++) {
      const int64_t a = q_indptr[q], b = q_indptr[q + 1];
      if (b < a || (b > a && (!q_ids || !q_tf))) return kv_fail(KV_ERR_INVALID, "kv_topk: bad query CSR");
      prep_query(ix, q_ids + a, q_tf + a, b - a, q_oov ? q_oov[q] : 0.0, qp[(size_t)q]);
      c_nq[q] = (float)qp[(size_t)q].nq;
      c_dotU[q] = (float)qp[(size_t)q].dotU;
      c_corrU[q] = (float)qp[(size_t)q].corrU;
      c_ninf[q] = -INFINITY;
      int nf = (int)qp[(size_t)q].fid.size(), nx = 0;
      for (uint32_t f : qp[(size_t)q].tfq) nx += f > 1;
      if (nf > H / 2 || nx > TXCAP) {  // too many features for a tile: full float64 scan + selection instead
        is_irr[(size_t)q] = 1;
        ix->irr_q.push_back(q);
        ix->irr_ids.insert(ix->irr_ids.end(), q_ids + a, q_ids + b);
        ix->irr_tf.insert(ix->irr_tf.end(), q_tf + a, q_tf + b);
        ix->irr_indptr.push_back((int64_t)ix->irr_ids.size());
        ix->irr_oov.push_back(q_oov ? q_oov[q] : 0.0);
        nf = 0; nx = 0;
      }
      if (cur.q_count == QT || cur_feats + nf > H / 2 || cur.n_extras + nx > TXCAP) {
        tiles.push_back(cur);
        cur = TileDesc{(int)q, 0, 0, 0};
        cur_feats = 0;
      }
      cur.q_count++;
      cur_feats += nf;
      cur.n_extras += nx;
    }
    tiles.push_back(cur);
  }
  const int64_t n_tiles = (int64_t)tiles.size();
  KV_CUDA(ix->h_tables.ensure(n_tiles * (int64_t)Tile::table_bytes));
  KV_CUDA(ix->h_tiles.ensure(n_tiles));
  for (int64_t t = 0; t < n_tiles; t++) {
    unsigned char *tb = ix->h_tables.p + (size_t)t * Tile::table_bytes;
    uint32_t *keys = (uint32_t *)(tb + Tile::off_keys);
    float *ad = (float *)(tb + Tile::off_ad);
    uint32_t *masks = (uint32_t *)(tb + Tile::off_masks);
    uint32_t *xkey = (uint32_t *)(tb + Tile::off_xkey);
    float *xtf = (float *)(tb + Tile::off_xtf);
    memset(keys, 0xFF, sizeof(uint32_t) * H);
    memset(ad, 0, Tile::table_bytes - Tile::off_ad);
    TileDesc &td = tiles[(size_t)t];
    int nx = 0;
    for (int qi = 0; qi < td.q_count; qi++) {
      int64_t q = td.q_begin + qi;
      if (is_irr[(size_t)q]) continue;
      const QueryPrep &p = qp[(size_t)q];
      for (size_t i = 0; i < p.fid.size(); i++) {
        uint32_t f = p.fid[i];
        uint32_t h = (f * 0x9E3779B1u) >> (32 - TLOGH);
        while (keys[h] != KEY_EMPTY && (keys[h] & FID_MASK) != f) h = (h + 1) & (H - 1);
        if (keys[h] == KEY_EMPTY) {
          keys[h] = f;
          double a, d;
          idf_host(ix->n_total, ix->h_df[f], a, d);
          ad[2 * h] = (float)a;
          ad[2 * h + 1] = (float)d;
        }
        masks[(size_t)h * TG + (qi >> 5)] |= 1u << (qi & 31);
        if (p.tfq[i] > 1) {
          keys[h] |= KEY_MULTI;
          xkey[nx] = (h << 8) | (uint32_t)qi;
          xtf[nx] = (float)p.tfq[i];
          nx++;
        }
      }
    }
    td.n_extras = nx;
    ix->h_tiles.p[t] = td;
  }
  KV_CUDA(ix->d_tables.ensure(n_tiles * (int64_t)Tile::table_bytes));
  KV_CUDA(ix->d_tiles.ensure(n_tiles));
  KV_CUDA(ix->d_qconst.ensure(4 * n_q));
  KV_CUDA(cudaEventRecord(ix->ev[0], s));
  KV_CUDA(cudaMemcpyAsync(ix->d_tables.p, ix->h_tables.p, (size_t)n_tiles * Tile::table_bytes, cudaMemcpyHostToDevice, s));
  KV_CUDA(cudaMemcpyAsync(ix->d_tiles.p, ix->h_tiles.p, (size_t)n_tiles * sizeof(TileDesc), cudaMemcpyHostToDevice, s));
  KV_CUDA(cudaMemcpyAsync(ix->d_qconst.p, ix->h_qconst.p, (size_t)4 * n_q * sizeof(float), cudaMemcpyHostToDevice, s));
  KV_CUDA(cudaEventRecord(ix->ev[1], s));
  KV_CUDA(cudaStreamSynchronize(s));  // the pinned staging buffers may be rewritten by the next call
  ix->batch_q = n_q;
  ix->batch_tiles = n_tiles;
  ix->batch_h2d_bytes = n_tiles * (int64_t)(Tile::table_bytes + sizeof(TileDesc)) + 4 * n_q * (int64_t)sizeof(float);
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
  // launch geometry: tiles x row-splits; aim at >= 8 waves of resident CTAs
  const int ctas_per_sm = 2;
  int64_t want = (int64_t)ix->sm_count * ctas_per_sm * 8;
  int64_t n_splits = (want + n_tiles - 1) / n_tiles;
  n_splits = std::max<int64_t>(1, std::min<int64_t>(n_splits, std::max<int64_t>(1, ix->n_chunks / 8)));
  n_splits = std::min<int64_t>(n_splits, 2048);
  ix->last_tiles = n_tiles; ix->last_splits = n_splits; ix->last_ctas = n_tiles * n_splits;
  KV_CUDA(ix->d_gthr.ensure(n_q));
  KV_CUDA(ix->d_part_s.ensure(n_splits * n_q * k));
  KV_CUDA(ix->d_part_r.ensure(n_splits * n_q * k));

  KV_CUDA(cudaEventRecord(ix->ev[1], s));
  if (ix->n_rows > 0) {
    // global lower bounds of the k-th score start at -inf (staged as the 4th constants column)
    KV_CUDA(cudaMemcpyAsync(ix->d_gthr.p, ix->d_qconst.p + 3 * n_q, (size_t)n_q * sizeof(float), cudaMemcpyDeviceToDevice, s));
    TopkParams P;
    P.stream = ix->d_stream.p; P.chunkptr = ix->d_chunkptr.p; P.n_chunks = ix->n_chunks; P.n_rows = ix->n_rows;
    P.row_base = ix->row_base; P.B32 = ix->d_B32.p; P.ovf_keys = ix->d_ovf_keys.p; P.ovf_vals = ix->d_ovf_vals.p;
    P.n_ovf = ix->n_ovf; P.tables = ix->d_tables.p; P.tiles = ix->d_tiles.p;
    P.q_nq = ix->d_qconst.p; P.q_dotU = P.q_nq + n_q; P.q_corrU = P.q_dotU + n_q;
    P.gthr = ix->d_gthr.p; P.n_q = n_q; P.k = k; P.n_splits = (int)n_splits;
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
                                                                         n_q, k, d_out_s, d_out_r);
    KV_CUDA(cudaGetLastError());
    for (size_t i = 0; i < ix->irr_q.size(); i++) {
      const int64_t q = ix->irr_q[i], a = ix->irr_indptr[i], b = ix->irr_indptr[i + 1];
      int rc = score_impl(ix, ix->irr_ids.data() + a, ix->irr_tf.data() + a, b - a, ix->irr_oov[i], nullptr);
      if (rc != KV_OK) return rc;
      select_topk_kernel<<<1, 1024, 0, s>>>(ix->d_scores.p, ix->n_rows, ix->row_base, k, d_out_s + q * k, d_out_r + q * k);
      KV_CUDA(cudaGetLastError());
    }
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
                                                                 n_lists, n_q, k, (float *)d_scores_out,
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

int kv_index_layout(const kv_index *ix, int64_t bytes[3], int64_t counts[8]) {
  if (!ix || !bytes || !counts) return kv_fail(KV_ERR_INVALID, "kv_index_layout: bad arguments");
  bytes[0] = ix->stream_len * 4;
  bytes[1] = ix->n_rows * 4;
  bytes[2] = (ix->n_chunks + 1) * 8;
  counts[0] = ix->stream_len; counts[1] = ix->n_univ; counts[2] = ix->n_rows;
  counts[3] = ix->last_ctas; counts[4] = ix->last_tiles; counts[5] = ix->last_splits;
  counts[6] = ix->batch_h2d_bytes; counts[7] = ix->n_ovf;
  return KV_OK;
}

}  // extern "C"
