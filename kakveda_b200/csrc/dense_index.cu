// K2: dense-embedding cosine scan with fused top-k on the 5th-gen tensor cores (BASELINE configs[1]:
// "1M-entry GFKB, 768-d embedding cosine, 10k-query batch").
//
// The reference has no embedding path (only TF-IDF exists, dense embeddings are a
// documented possible extension, docs/failure-intelligence.md:43-46) -- parity for this kernel is
// UNPINNED; its oracle is oracle/tfidf_oracle.py::dense_cosine (float64 on the same bf16 inputs).
//
// scores[q, r] = <Q[q,:], C[r,:]> / (|Q[q]| |C[r]|): a bf16 GEMM Q * C^T with fp32 accumulation, the
// norms applied as fp32 scales in the epilogue, and the top-k fused into the epilogue so that the
// [Q, N] score matrix never exists.  A work item is one (128-query tile, row split); items are ordered
// split-major so that the CTAs resident together (one per SM) stream the same corpus rows through L2.  Few, long
// row splits keep every list's k-th-score threshold high.  Per CTA (320 threads):
//   warp 0      TMA producer: K-slices (64 elements) of the query tile and of the row tile -> 3-stage
//               shared-memory ring (128B-swizzled), mbarrier expect_tx / complete_tx
//   warp 1      MMA issuer: tcgen05.mma cta_group::1 kind::f16, M=128 N=256 K=16, accumulators in TMEM
//               (2 x 256 columns, double buffered); tcgen05.commit releases ring slots / publishes a tile
//   warps 2-9   epilogue (two warps per TMEM lane quarter, each scanning half of the 256 columns): tcgen05.ld of
//               the thread's TMEM lane (= its query), scale, threshold test, insertion into the thread's own sorted
//               top-k list (k <= 32; ties keep the lower row; a self-join skips the query's own row)
// Each (query tile, split, column half) writes one partial list; K5 (kv_merge_topk_device) merges them.  CTAs
// working on the same queries exchange k-th-score lower bounds through global memory (gthr).
#include "kv_cuda.cuh"

#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace {

constexpr int BM = 128;        // queries per CTA (TMEM lanes)
constexpr int BN = 256;        // corpus rows per MMA tile (TMEM columns per accumulator stage)
constexpr int BK = 64;         // K-slice: 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 3;       // 3 x 48 KiB ring + two list sets fit in 227 KiB
constexpr int A_BYTES = BM * BK * 2;  // 16 KiB
constexpr int B_BYTES = BN * BK * 2;  // 32 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int MAXK = 32;       // per-query list slots: 2 x 32 x 128 x 8 B = 64 KiB beside the 3 x 48 KiB ring (k <= 32)
constexpr int N_THREADS = 320;  // 10 warps: TMA, MMA, 8 epilogue
constexpr int EPI_THREADS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc(const void *smem) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_u32(smem) & 0x3FFFF) >> 4);  // start address
  d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset
  d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}

// instruction descriptor: D=f32, A=B=bf16, both K-major, N=256, M=128
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct DenseParams {
  int64_t n_rows, row_base, n_q;
  int64_t excl_base;        // self-join: query q must not match global row excl_base + q (-1: no exclusion)
  int64_t r_tiles, q_tiles;  // 256-row tiles, 128-query tiles
  int dim, k, n_lists, dbg;
  const float *inv_norm_c;  // [n_rows]
  const float *inv_norm_q;  // [n_q]
  unsigned int *gthr;       // [n_q] order-preserving key of a lower bound of the global k-th score (0 = none)
  float *part_scores;       // [n_lists][n_q][k]
  long long *part_rows;
};

// order-preserving float <-> unsigned key (cosines may be negative)
__device__ __forceinline__ unsigned int fkey(float f) {
  unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

struct __align__(1024) DenseSmem {
  unsigned char stage[STAGES][STAGE_BYTES];
  uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
  __align__(16) float inv_c[2][BN];
  float lscore[2][MAXK][BM];  // [column half][slot][query]: the 32 lanes of a warp hit 32 different banks
  int lrow[2][MAXK][BM];
};

__global__ void __launch_bounds__(N_THREADS, 1)
dense_topk_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_c, DenseParams P) {
  extern __shared__ unsigned char smem_raw[];
  // aligned by an OFFSET into the shared array (not by rounding a generic pointer): the compiler keeps the shared
  // address space, so list and norm accesses are LDS/STS instead of generic loads / stores
  DenseSmem &S = *reinterpret_cast<DenseSmem *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work item = (query tile, row split); consecutive CTAs take consecutive query tiles of the same split, so the
  // CTAs resident at one time stream the same corpus rows (B tiles are shared through L2)
  const int64_t item = blockIdx.x;
  const int64_t my_qtile = item % P.q_tiles, my_split = item / P.q_tiles;
  const int64_t L0 = my_qtile * P.r_tiles + P.r_tiles * my_split / P.n_lists;
  const int64_t L1 = my_qtile * P.r_tiles + P.r_tiles * (my_split + 1) / P.n_lists;
  const int n_kb = P.dim / BK;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < STAGES; i++) { mbar_init(&S.full_bar[i], 1); mbar_init(&S.empty_bar[i], 1); }
    for (int i = 0; i < 2; i++) { mbar_init(&S.tmem_full[i], 1); mbar_init(&S.tmem_empty[i], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: 512 columns = two 128x256 fp32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&S.tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = S.tmem_base;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t L = L0; L < L1; L++) {
        const int qtile = (int)(L / P.r_tiles);
        const int64_t t = L % P.r_tiles;
        for (int kb = 0; kb < n_kb; kb++) {
          mbar_wait(&S.empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&S.full_bar[stage], STAGE_BYTES);
          tma_load_2d(S.stage[stage], &map_q, &S.full_bar[stage], kb * BK, qtile * BM);
          tma_load_2d(S.stage[stage] + A_BYTES, &map_c, &S.full_bar[stage], kb * BK, (int)(t * BN));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int64_t it = 0;
      for (int64_t L = L0; L < L1; L++, it++) {
        const int as = (int)(it & 1);
        const uint32_t aphase = (uint32_t)((it >> 1) & 1);
        mbar_wait(&S.tmem_empty[as], aphase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < n_kb; kb++) {
          mbar_wait(&S.full_bar[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = umma_desc(S.stage[stage]);
          const uint64_t db = umma_desc(S.stage[stage] + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; k++)  // advance 32 bytes (2 x 16-byte units) per K=16 step inside the swizzle row
            umma_f16(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), (uint32_t)((kb | k) != 0));
          umma_commit(&S.empty_bar[stage]);  // slot free once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&S.tmem_full[as]);  // accumulator complete
      }
    }
  } else {
    // ===== epilogue: warps 2..9; a warp may only touch TMEM lanes 32*(warp%4) .. +31, so two warps share
    // each lane quarter and split the 256 columns of a tile between them (half 0 / half 1) =====
    const int lane_base = 32 * (warp & 3);
    const int qi = lane_base + lane;            // TMEM lane = query inside the tile
    const int et = (warp - 2) * 32 + lane;      // 0..255 among the epilogue threads
    const int half = (warp - 2) >> 2;           // which 128 columns of every tile this thread scans
    const int k = P.k;
    float *ls = &S.lscore[half][0][qi];         // element j of this thread's list lives at ls[j * BM]
    int *lr = &S.lrow[half][0][qi];
    int cur_qtile = -1, cnt = 0;
    int excl = -1;          // local row this thread's query must not match (self-join)
    int64_t q = 0;
    bool q_ok = false;
    float inv_q = 0.f;
    float thr = -INFINITY;  // own k-th score: later rows must beat it strictly (rows ascend inside a CTA)
    float gth = -INFINITY;  // k-th score another CTA already secured for this query: ties may still win on row id
    float gth_pred = -INFINITY;  // largest float below gth
    float lo = INFINITY;         // a row enters the list iff its score > lo
    auto flush = [&]() {    // publish the list of (cur_qtile, this CTA)
      if (cur_qtile < 0 || !q_ok) return;
      const int slot = (int)my_split * 2 + half;
      for (int j = 0; j < k; j++) {
        const size_t o = ((size_t)slot * P.n_q + q) * k + j;
        P.part_scores[o] = j < cnt ? ls[j * BM] : -INFINITY;
        P.part_rows[o] = j < cnt ? (long long)(P.row_base + lr[j * BM]) : -1LL;
      }
    };
    int64_t it = 0;
    for (int64_t L = L0; L < L1; L++, it++) {
      const int qtile = (int)(L / P.r_tiles);
      const int64_t t = L % P.r_tiles;
      if (qtile != cur_qtile) {  // (warp-uniform) next query tile: publish and restart the lists
        flush();
        cur_qtile = qtile;
        q = (int64_t)qtile * BM + qi;
        q_ok = q < P.n_q;
        inv_q = q_ok ? P.inv_norm_q[q] : 0.f;
        {
          const int64_t e = P.excl_base >= 0 ? P.excl_base + q - P.row_base : -1;
          excl = (e >= 0 && e < P.n_rows) ? (int)e : -1;
        }
        cnt = 0;
        thr = gth = gth_pred = -INFINITY;
        lo = q_ok ? -INFINITY : INFINITY;
        for (int j = 0; j < k; j++) { ls[j * BM] = -INFINITY; lr[j * BM] = 0x7fffffff; }
      }
      const int as = (int)(it & 1);
      const uint32_t aphase = (uint32_t)((it >> 1) & 1);
      // inverse norms of this tile's rows (0 past the end; such rows are rejected by index below)
      const int64_t row0 = t * BN;
      for (int c = et; c < BN; c += EPI_THREADS) S.inv_c[as][c] = (row0 + c < P.n_rows) ? P.inv_norm_c[row0 + c] : -INFINITY;  // 0 * -inf = NaN: never a candidate
      if (q_ok) {
        const unsigned int gk = *(volatile unsigned int *)&P.gthr[q];
        if (gk > fkey(-INFINITY) && fkey_inv(gk) > gth) {
          gth = fkey_inv(gk);
          gth_pred = fkey_inv(gk - 1);  // the order-preserving key makes "previous float" a decrement
          lo = fmaxf(thr, gth_pred);
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&S.tmem_full[as], aphase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(as * BN);
#pragma unroll 1
      for (int c0 = half * (BN / 2); c0 < ((P.dbg == 1) ? 0 : (half + 1) * (BN / 2)); c0 += 32) {
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr + (uint32_t)c0)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (P.dbg == 2) { if (v[0] == 0x12345678u && v[31] == 0x9abcdef0u) thr = 1.f; continue; }
        // 32 independent scale ops + a max tree: one (rarely taken) branch per 32 rows instead of 32
        float tv[32];
        const float4 *ic4 = reinterpret_cast<const float4 *>(&S.inv_c[as][c0]);
#pragma unroll
        for (int j4 = 0; j4 < 8; j4++) {
          const float4 ic = ic4[j4];
          tv[4 * j4 + 0] = __uint_as_float(v[4 * j4 + 0]) * ic.x;
          tv[4 * j4 + 1] = __uint_as_float(v[4 * j4 + 1]) * ic.y;
          tv[4 * j4 + 2] = __uint_as_float(v[4 * j4 + 2]) * ic.z;
          tv[4 * j4 + 3] = __uint_as_float(v[4 * j4 + 3]) * ic.w;
        }
        // per 8 columns: max -> one branch; only sub-blocks where some query of the warp has a candidate are
        // walked element by element (`sc > lo` == `sc > thr && sc >= gth`, lo = max(thr, pred(gth)))
#pragma unroll
        for (int sb = 0; sb < 4; sb++) {
          const float m01 = fmaxf(tv[8 * sb + 0], tv[8 * sb + 1]), m23 = fmaxf(tv[8 * sb + 2], tv[8 * sb + 3]);
          const float m45 = fmaxf(tv[8 * sb + 4], tv[8 * sb + 5]), m67 = fmaxf(tv[8 * sb + 6], tv[8 * sb + 7]);
          const float best = fmaxf(fmaxf(m01, m23), fmaxf(m45, m67)) * inv_q;
          if (best > lo) {
#pragma unroll  // static indices keep tv[] in registers
            for (int j = 8 * sb; j < 8 * sb + 8; j++) {
              const float sc = tv[j] * inv_q;
              if (sc > lo && (int)(row0 + c0 + j) != excl) {
                int pos = cnt < k ? cnt++ : k - 1;
                while (pos > 0 && ls[(pos - 1) * BM] < sc) {
                  ls[pos * BM] = ls[(pos - 1) * BM];
                  lr[pos * BM] = lr[(pos - 1) * BM];
                  pos--;
                }
                ls[pos * BM] = sc;
                lr[pos * BM] = (int)(row0 + c0 + j);
                if (cnt == k) { thr = ls[(k - 1) * BM]; lo = fmaxf(thr, gth_pred); }
              }
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.tmem_empty[as]);
      if (q_ok && cnt == k && thr > gth) atomicMax(&P.gthr[q], fkey(thr));
    }
    flush();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// inverse L2 norms of bf16 rows (float64 accumulation); zero rows get 0 (their scores are 0)
__global__ void inv_norm_kernel(const __nv_bfloat16 *__restrict__ x, int64_t n, int dim, float *out) {
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n) return;
  double s = 0.0;
  for (int i = lane; i < dim; i += 32) {
    double v = (double)__bfloat162float(x[r * dim + i]);
    s += v * v;
  }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
  if (lane == 0) out[r] = s > 0.0 ? (float)(1.0 / sqrt(s)) : 0.f;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(CUtensorMap *map, const void *base, int64_t rows, int dim, int box_rows) {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p)
      return kv_fail(KV_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    fn = (PFN_encodeTiled)p;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)dim * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return kv_fail(KV_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return KV_OK;
}

}  // namespace

struct kv_dense_index {
  int device = 0, dim = 0;
  int64_t row_base = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  std::mutex mu;
  int sm_count = 148;
  DevVec<__nv_bfloat16> rows;
  int64_t n_rows = 0;
  DevBuf<float> d_inv_c, d_inv_q, d_part_s, d_out_s;
  DevBuf<long long> d_part_r, d_out_r;
  DevBuf<__nv_bfloat16> d_q;
  DevBuf<unsigned int> d_gthr;
  bool finalized = false;
  float last_ms = 0;
  int64_t last_splits = 0;
};

extern "C" {

int kv_dense_create(int device, int dim, int64_t row_base, kv_dense_index **out) {
  if (!out) return kv_fail(KV_ERR_INVALID, "kv_dense_create: out is NULL");
  if (dim < BK || dim % BK != 0 || dim > 8192) return kv_fail(KV_ERR_INVALID, "kv_dense_create: dim must be a multiple of 64 (64..8192)");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return kv_fail(KV_ERR_CUDA, "kv_dense_create: no CUDA device visible (this library has no CPU path)");
  }
  if (device < 0 || device >= n) return kv_fail(KV_ERR_INVALID, "kv_dense_create: device %d out of range", device);
  KV_CUDA(cudaSetDevice(device));
  kv_dense_index *dx = new kv_dense_index();
  dx->device = device;
  dx->dim = dim;
  dx->row_base = row_base;
  cudaDeviceProp prop;
  KV_CUDA(cudaGetDeviceProperties(&prop, device));
  dx->sm_count = prop.multiProcessorCount;
  KV_CUDA(cudaStreamCreateWithFlags(&dx->stream, cudaStreamNonBlocking));
  for (auto &e : dx->ev) KV_CUDA(cudaEventCreate(&e));
  KV_CUDA(cudaFuncSetAttribute(dense_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DenseSmem) + 1024));
  *out = dx;
  return KV_OK;
}

void kv_dense_destroy(kv_dense_index *dx) {
  if (!dx) return;
  cudaSetDevice(dx->device);
  cudaStreamSynchronize(dx->stream);
  dx->rows.release(); dx->d_inv_c.release(); dx->d_inv_q.release(); dx->d_part_s.release(); dx->d_out_s.release();
  dx->d_part_r.release(); dx->d_out_r.release(); dx->d_q.release(); dx->d_gthr.release();
  for (auto &e : dx->ev) if (e) cudaEventDestroy(e);
  if (dx->stream) cudaStreamDestroy(dx->stream);
  delete dx;
}

int64_t kv_dense_rows(const kv_dense_index *dx) { return dx ? dx->n_rows : 0; }

// rows: n x dim bfloat16 (raw uint16 bit patterns), row-major, host memory
int kv_dense_append(kv_dense_index *dx, const uint16_t *rows_bf16, int64_t n) {
  if (!dx || n < 0 || (n > 0 && !rows_bf16)) return kv_fail(KV_ERR_INVALID, "kv_dense_append: bad arguments");
  if (n == 0) return KV_OK;
  std::lock_guard<std::mutex> g(dx->mu);
  KV_CUDA(cudaSetDevice(dx->device));
  if (dx->n_rows + n >= (1LL << 31) - BN) return kv_fail(KV_ERR_INVALID, "kv_dense_append: more than 2^31 rows in one shard");
  KV_CUDA(dx->rows.reserve((dx->n_rows + n) * dx->dim, dx->stream));
  KV_CUDA(cudaMemcpyAsync(dx->rows.p + dx->n_rows * dx->dim, rows_bf16, (size_t)n * dx->dim * 2, cudaMemcpyHostToDevice, dx->stream));
  KV_CUDA(cudaStreamSynchronize(dx->stream));
  dx->n_rows += n;
  dx->rows.n = dx->n_rows * dx->dim;
  dx->finalized = false;
  return KV_OK;
}

int kv_dense_finalize(kv_dense_index *dx) {
  if (!dx) return kv_fail(KV_ERR_INVALID, "kv_dense_finalize: NULL handle");
  std::lock_guard<std::mutex> g(dx->mu);
  KV_CUDA(cudaSetDevice(dx->device));
  KV_CUDA(dx->d_inv_c.ensure(std::max<int64_t>(dx->n_rows, 1)));
  if (dx->n_rows) {
    inv_norm_kernel<<<(unsigned)((dx->n_rows * 32 + 255) / 256), 256, 0, dx->stream>>>(dx->rows.p, dx->n_rows, dx->dim, dx->d_inv_c.p);
    KV_CUDA(cudaGetLastError());
  }
  KV_CUDA(cudaStreamSynchronize(dx->stream));
  dx->finalized = true;
  return KV_OK;
}

// scan + merge of n_q queries already on the device (d_q: bf16 [n_q, dim], 16-byte aligned) into device buffers
static int dense_run(kv_dense_index *dx, const __nv_bfloat16 *d_q, int64_t n_q, int k, int64_t excl_base, float *d_out_s,
                     long long *d_out_r) {
  cudaStream_t s = dx->stream;
  if (dx->n_rows == 0) {
    KV_CUDA(cudaMemsetAsync(d_out_r, 0xFF, (size_t)n_q * k * 8, s));  // row -1
    std::vector<float> neg((size_t)(n_q * k), -INFINITY);
    KV_CUDA(cudaMemcpyAsync(d_out_s, neg.data(), neg.size() * 4, cudaMemcpyHostToDevice, s));
    KV_CUDA(cudaStreamSynchronize(s));
    return KV_OK;
  }
  KV_CUDA(dx->d_inv_q.ensure(n_q));
  inv_norm_kernel<<<(unsigned)((n_q * 32 + 255) / 256), 256, 0, s>>>(d_q, n_q, dx->dim, dx->d_inv_q.p);
  KV_CUDA(cudaGetLastError());
  CUtensorMap map_q, map_c;
  int rc = make_map(&map_q, d_q, n_q, dx->dim, BM);
  if (rc != KV_OK) return rc;
  rc = make_map(&map_c, dx->rows.p, dx->n_rows, dx->dim, BN);
  if (rc != KV_OK) return rc;
  const int64_t q_tiles = (n_q + BM - 1) / BM, r_tiles = (dx->n_rows + BN - 1) / BN;
  // row splits: as few as possible (long row ranges keep the k-th-score thresholds high) while the CTA count
  // fills whole waves of SMs (one CTA per SM)
  int64_t n_lists = 1;
  {
    double best = 1e18;
    const int64_t lo = std::max<int64_t>(1, (dx->sm_count + q_tiles - 1) / q_tiles);
    for (int64_t sp = lo; sp <= std::min<int64_t>(r_tiles, lo + 16); sp++) {
      const int64_t ctas = q_tiles * sp, waves = (ctas + dx->sm_count - 1) / dx->sm_count;
      const double cost = (double)(waves * dx->sm_count) / (double)ctas * (1.0 + 0.01 * (double)sp);
      if (cost < best) { best = cost; n_lists = sp; }
    }
    n_lists = std::max<int64_t>(1, std::min<int64_t>(n_lists, r_tiles));
  }
  dx->last_splits = n_lists;
  const int64_t n_part = n_lists * 2;  // two epilogue threads (column halves) per query and CTA
  KV_CUDA(dx->d_part_s.ensure(n_part * n_q * k)); KV_CUDA(dx->d_part_r.ensure(n_part * n_q * k));
  KV_CUDA(dx->d_gthr.ensure(n_q));
  KV_CUDA(cudaMemsetAsync(dx->d_gthr.p, 0, (size_t)n_q * 4, s));
  // unused (query tile, slot) pairs stay "empty": row -1 (the merge ignores their scores)
  KV_CUDA(cudaMemsetAsync(dx->d_part_r.p, 0xFF, (size_t)n_part * n_q * k * 8, s));
  KV_CUDA(cudaMemsetAsync(dx->d_part_s.p, 0xFF, (size_t)n_part * n_q * k * 4, s));
  DenseParams P;
  P.n_rows = dx->n_rows; P.row_base = dx->row_base; P.n_q = n_q; P.dim = dx->dim; P.k = k; P.n_lists = (int)n_lists;
  P.excl_base = excl_base;
  P.r_tiles = r_tiles; P.q_tiles = q_tiles;
  P.dbg = getenv("KAKVEDA_B200_DENSE_DBG") ? atoi(getenv("KAKVEDA_B200_DENSE_DBG")) : 0;
  P.inv_norm_c = dx->d_inv_c.p; P.inv_norm_q = dx->d_inv_q.p; P.gthr = dx->d_gthr.p;
  P.part_scores = dx->d_part_s.p; P.part_rows = dx->d_part_r.p;
  const int64_t grid = q_tiles * n_lists;
  KV_CUDA(cudaEventRecord(dx->ev[0], s));
  dense_topk_kernel<<<(unsigned)grid, N_THREADS, sizeof(DenseSmem) + 1024, s>>>(map_q, map_c, P);
  KV_CUDA(cudaGetLastError());
  KV_CUDA(cudaEventRecord(dx->ev[1], s));
  KV_CUDA(cudaStreamSynchronize(s));
  cudaEventElapsedTime(&dx->last_ms, dx->ev[0], dx->ev[1]);
  return kv_merge_topk_device(dx->device, dx->d_part_s.p, dx->d_part_r.p, (int)n_part, n_q, k, d_out_s, d_out_r);
}

// q: n_q x dim bfloat16 bit patterns (host).  Outputs (host): scores float32[n_q*k], rows int64[n_q*k],
// ordered by (score desc, row asc); unused slots (-inf, -1).
int kv_dense_topk(kv_dense_index *dx, const uint16_t *q_bf16, int64_t n_q, int k, float *out_scores, int64_t *out_rows) {
  if (!dx || n_q < 0 || k < 1 || k > MAXK || (n_q > 0 && (!q_bf16 || !out_scores || !out_rows)))
    return kv_fail(KV_ERR_INVALID, "kv_dense_topk: bad arguments (k must be 1..32)");
  std::lock_guard<std::mutex> g(dx->mu);
  if (!dx->finalized) return kv_fail(KV_ERR_STATE, "kv_dense_topk: index not finalized");
  if (n_q == 0) return KV_OK;
  KV_CUDA(cudaSetDevice(dx->device));
  cudaStream_t s = dx->stream;
  KV_CUDA(dx->d_q.ensure(n_q * dx->dim));
  KV_CUDA(cudaMemcpyAsync(dx->d_q.p, q_bf16, (size_t)n_q * dx->dim * 2, cudaMemcpyHostToDevice, s));
  KV_CUDA(dx->d_out_s.ensure(n_q * k)); KV_CUDA(dx->d_out_r.ensure(n_q * k));
  int rc = dense_run(dx, dx->d_q.p, n_q, k, -1, dx->d_out_s.p, dx->d_out_r.p);
  if (rc != KV_OK) return rc;
  KV_CUDA(cudaMemcpy(out_scores, dx->d_out_s.p, (size_t)n_q * k * 4, cudaMemcpyDeviceToHost));
  KV_CUDA(cudaMemcpy(out_rows, dx->d_out_r.p, (size_t)n_q * k * 8, cudaMemcpyDeviceToHost));
  return KV_OK;
}

int kv_dense_topk_device(kv_dense_index *dx, const void *d_q_bf16, int64_t n_q, int k, int64_t exclude_base, void *d_scores,
                         void *d_rows) {
  if (!dx || n_q < 0 || k < 1 || k > MAXK || (n_q > 0 && (!d_q_bf16 || !d_scores || !d_rows)))
    return kv_fail(KV_ERR_INVALID, "kv_dense_topk_device: bad arguments (k must be 1..32)");
  if (((uintptr_t)d_q_bf16 & 15) != 0) return kv_fail(KV_ERR_INVALID, "kv_dense_topk_device: queries must be 16-byte aligned");
  std::lock_guard<std::mutex> g(dx->mu);
  if (!dx->finalized) return kv_fail(KV_ERR_STATE, "kv_dense_topk_device: index not finalized");
  if (n_q == 0) return KV_OK;
  KV_CUDA(cudaSetDevice(dx->device));
  return dense_run(dx, (const __nv_bfloat16 *)d_q_bf16, n_q, k, exclude_base, (float *)d_scores, (long long *)d_rows);
}

// rows already on the device (e.g. a torch tensor): device-to-device append
int kv_dense_append_device(kv_dense_index *dx, const void *d_rows_bf16, int64_t n) {
  if (!dx || n < 0 || (n > 0 && !d_rows_bf16)) return kv_fail(KV_ERR_INVALID, "kv_dense_append_device: bad arguments");
  if (n == 0) return KV_OK;
  std::lock_guard<std::mutex> g(dx->mu);
  KV_CUDA(cudaSetDevice(dx->device));
  if (dx->n_rows + n >= (1LL << 31) - BN) return kv_fail(KV_ERR_INVALID, "kv_dense_append_device: more than 2^31 rows in one shard");
  KV_CUDA(dx->rows.reserve((dx->n_rows + n) * dx->dim, dx->stream));
  KV_CUDA(cudaMemcpyAsync(dx->rows.p + dx->n_rows * dx->dim, d_rows_bf16, (size_t)n * dx->dim * 2, cudaMemcpyDeviceToDevice, dx->stream));
  KV_CUDA(cudaStreamSynchronize(dx->stream));
  dx->n_rows += n;
  dx->rows.n = dx->n_rows * dx->dim;
  dx->finalized = false;
  return KV_OK;
}

// all-pairs (BASELINE configs[3]): local rows [q_begin, q_end) as queries against the whole shard, each row's own
// entry excluded; outputs on the device
int kv_dense_selfjoin_device(kv_dense_index *dx, int64_t q_begin, int64_t q_end, int k, void *d_scores, void *d_rows) {
  if (!dx || q_begin < 0 || q_end < q_begin || k < 1 || k > MAXK || !d_scores || !d_rows)
    return kv_fail(KV_ERR_INVALID, "kv_dense_selfjoin_device: bad arguments (k must be 1..32)");
  std::lock_guard<std::mutex> g(dx->mu);
  if (!dx->finalized) return kv_fail(KV_ERR_STATE, "kv_dense_selfjoin_device: index not finalized");
  if (q_end > dx->n_rows) return kv_fail(KV_ERR_INVALID, "kv_dense_selfjoin_device: row range outside the index");
  if (q_end == q_begin) return KV_OK;
  KV_CUDA(cudaSetDevice(dx->device));
  return dense_run(dx, dx->rows.p + q_begin * dx->dim, q_end - q_begin, k, dx->row_base + q_begin, (float *)d_scores,
                   (long long *)d_rows);
}

int kv_dense_last_timing(const kv_dense_index *dx, float *gemm_ms, int64_t *splits) {
  if (!dx || !gemm_ms || !splits) return kv_fail(KV_ERR_INVALID, "kv_dense_last_timing: bad arguments");
  *gemm_ms = dx->last_ms;
  *splits = dx->last_splits;
  return KV_OK;
}

}  // extern "C"
