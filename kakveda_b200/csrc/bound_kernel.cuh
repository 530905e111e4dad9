// K1b-B: upper bounds of every (query, chunk) score on the 5th-gen tensor cores, and the candidate lists they leave.
//
// For a query q and a chunk c (32 rows), with U_c(t) = largest tf of feature t in the chunk (0 if absent) and
// Bmin_c = smallest positive row norm of the chunk:
//     dot(q, r) <= dotS_q + dotX_q + sum_{t frequent} W_q(t) U_c(t) + sum_{t rare, in the chunk} w_tile(t) U_c(t)
//     B_r + corr(q, r) >= Bmin_c + corrS_q                                  (corrS_q: every d(t) of q at its largest tf)
// so  score(q, r) <= ub(q, c) = dot_bound / sqrt(|q|^2 (Bmin_c + corrS_q))  for every row r of c  (exact block-max
// pruning: a chunk whose bound is below a query's k-th best score cannot hold a top-k row for it).
//   * the sum over the NF = 256 FREQUENT features is a [128 queries x 256] x [256 x 128 chunks] fp16 GEMM (weights
//     rounded UP to fp16, tf exact): tcgen05.mma cta_group::1 kind::f16, M = N = 128, fp32 accumulators in TMEM
//     (double buffered), the query operand resident in shared memory for the CTA's life, the chunk operand streamed
//     by TMA (128B swizzle, mbarrier expect_tx) -- this only computes BOUNDS; scores stay exact integer sums (K1b-S);
//   * the sum over RARE features is a join: the worker warps probe each chunk's rare block entries in the tile's
//     rare-feature table (shared memory) and add the hits into R[chunk][query] (shared-memory atomics);
//   * epilogue (16 warps, thread = query, tcgen05.ld of 32 chunk columns): bound vs the query's threshold.
//     pass 0 keeps, per thread, the 4 best chunks by bound (seeds: K1b-S scores them first, which gives every query a
//     close lower bound theta0 of its k-th best score); pass 1 appends {chunk, mask of the group's surviving queries}
//     to the scan group's candidate list (paged pool) for every chunk with ub >= theta0.
// One CTA = one 128-query tile x one range of 128-chunk blocks; 18 warps: 16 workers (join + epilogue), TMA, MMA.
#pragma once
#include "tfidf_kernels.cuh"

#include <cuda.h>

namespace kvk {

constexpr int B_BN = 128;                    // chunks per block = N of the MMA tile
constexpr int B_BK = 64;                     // K slice: 64 fp16 = one 128-byte swizzle row
constexpr int B_KSLICES = NF / B_BK;         // 4
constexpr int B_STAGES = 2;
constexpr int B_SLICE_BYTES = 128 * B_BK * 2;  // 16 KiB: one K slice of either operand
constexpr int B_WORKERS = 16;
constexpr int B_THREADS = (B_WORKERS + 2) * 32;
constexpr int B_SEEDS = 4;                   // seeds per worker thread; a query is served by 4 threads (column quarters)
constexpr int B_SEEDS_PER_QUERY = 4 * B_SEEDS;

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_addr(dst)),
      "l"((uint64_t)map), "r"(smem_addr(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw128(const void *smem) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr(smem) & 0x3FFFF) >> 4);  // start address
  d |= (uint64_t)1 << 16;                              // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset
  d |= (uint64_t)1 << 46;                              // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
  return d;
}

// instruction descriptor: D = f32, A = B = f16, both K-major, N = 128, M = 128
constexpr uint32_t B_IDESC = (1u << 4) | ((uint32_t)(B_BN >> 3) << 17) | ((uint32_t)(TILE_Q >> 4) << 24);

__device__ __forceinline__ void umma_f16_128(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(B_IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_1(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}

struct BoundParams {
  const uint32_t *blk;
  const BlockInfo *binfo;
  const float *chunk_minB;  // [n_chunks_pad]
  const unsigned long long *ovf_keys;
  const uint32_t *ovf_vals;
  int n_ovf;
  int64_t n_chunks, n_q;
  const unsigned char *rtab;                              // [n_tiles][RTAB_BYTES]
  const float *q_nq, *q_dotS, *q_dotX, *q_corrS;          // [n_q] (sorted query order)
  const int *gthr;                                        // [n_q] float bits: lower bound of the k-th score
  int pass;                                               // 0: seeds, 1: candidate lists
  int n_bsplits, jaccard;
  int *seeds;                                             // pass 0: [n_q][n_bsplits][B_SEEDS_PER_QUERY] chunk ids, -1 = none
  // pass 1: list l = group * n_bsplits + bsplit
  uint32_t *list_count;   // [n_lists]
  uint32_t *list_pages;   // [n_lists][max_pages]
  int max_pages;
  uint2 *pool;
  unsigned int *pool_next;  // pages handed out
  unsigned int pool_pages;  // capacity
  int *overflow;            // set when the pool ran out (the batch is rerun with a larger pool)
  unsigned long long *stats;  // [2] surviving (query, chunk) pairs, [3] candidate records
};

struct __align__(1024) BoundSmem {
  unsigned char a[B_KSLICES][B_SLICE_BYTES];   // the tile's weight rows, resident
  unsigned char b[B_STAGES][B_SLICE_BYTES];    // chunk slices in flight
  float R[B_BN][TILE_Q];                       // rare part of the dot bound, [chunk][query]
  unsigned char rtab[RTAB_BYTES];
  float minB[2][B_BN];
  uint64_t full_bar[B_STAGES], empty_bar[B_STAGES], a_bar, tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
  unsigned int lcount[4];
  int pages[1];  // [4][max_pages], sized at launch
};

static inline size_t bound_smem_bytes(int max_pages) { return sizeof(BoundSmem) + (size_t)4 * max_pages * sizeof(int) + 1024; }

__global__ void __launch_bounds__(B_THREADS, 1)
tfidf_bound_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_u, BoundParams P) {
  extern __shared__ unsigned char smem_raw[];
  BoundSmem &S = *reinterpret_cast<BoundSmem *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, bsplit = blockIdx.y;
  const int64_t n_blocks = (P.n_chunks + B_BN - 1) / B_BN;
  const int64_t blk_lo = n_blocks * bsplit / P.n_bsplits, blk_hi = n_blocks * (bsplit + 1) / P.n_bsplits;

  if (threadIdx.x == 0) {
    for (int i = 0; i < B_STAGES; i++) { mbar_init(&S.full_bar[i], 1); mbar_init(&S.empty_bar[i], 1); }
    mbar_init(&S.a_bar, 1);
    for (int i = 0; i < 2; i++) { mbar_init(&S.tmem_full[i], 1); mbar_init(&S.tmem_empty[i], B_WORKERS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == B_WORKERS + 1) {  // TMEM: 256 columns = two 128x128 fp32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_addr(&S.tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {  // rare-feature table of the tile, R = 0, list state
    const uint4 *src = (const uint4 *)(P.rtab + (size_t)tile * RTAB_BYTES);
    uint4 *dst = (uint4 *)S.rtab;
    for (int i = threadIdx.x; i < RTAB_BYTES / 16; i += B_THREADS) dst[i] = src[i];
    float4 *r4 = (float4 *)&S.R[0][0];
    for (int i = threadIdx.x; i < B_BN * TILE_Q / 4; i += B_THREADS) r4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = threadIdx.x; i < 4 * P.max_pages; i += B_THREADS) S.pages[i] = -1;
    if (threadIdx.x < 4) S.lcount[threadIdx.x] = 0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = S.tmem_base;

  if (warp == B_WORKERS) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_expect_tx(&S.a_bar, B_KSLICES * B_SLICE_BYTES);
      for (int s = 0; s < B_KSLICES; s++) tma_load_2d(S.a[s], &map_w, &S.a_bar, s * B_BK, tile * TILE_Q);
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t bk = blk_lo; bk < blk_hi; bk++) {
        for (int s = 0; s < B_KSLICES; s++) {
          mbar_wait(&S.empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&S.full_bar[stage], B_SLICE_BYTES);
          tma_load_2d(S.b[stage], &map_u, &S.full_bar[stage], s * B_BK, (int)(bk * B_BN));
          if (++stage == B_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == B_WORKERS + 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      mbar_wait(&S.a_bar, 0);
      int stage = 0;
      uint32_t phase = 0;
      int64_t it = 0;
      for (int64_t bk = blk_lo; bk < blk_hi; bk++, it++) {
        const int as = (int)(it & 1);
        const uint32_t aphase = (uint32_t)((it >> 1) & 1);
        mbar_wait(&S.tmem_empty[as], aphase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * B_BN);
        for (int s = 0; s < B_KSLICES; s++) {
          mbar_wait(&S.full_bar[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = umma_desc_sw128(S.a[s]);
          const uint64_t db = umma_desc_sw128(S.b[stage]);
#pragma unroll
          for (int kk = 0; kk < B_BK / 16; kk++)  // advance 32 bytes (2 x 16-byte units) per K=16 step inside the swizzle row
            umma_f16_128(tmem_d, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), (uint32_t)((s | kk) != 0));
          umma_commit_1(&S.empty_bar[stage]);  // slot free once these MMAs have read it
          if (++stage == B_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_1(&S.tmem_full[as]);  // accumulator complete
      }
    }
  } else {
    // ===== workers: join, then epilogue, per block =====
    const int qtr = warp & 3;   // TMEM lane quarter = scan group of the tile
    const int cs = warp >> 2;   // which 32 chunk columns of a block this warp's epilogue covers
    const int qi = qtr * 32 + lane;
    const int64_t slot = (int64_t)tile * TILE_Q + qi;
    const bool q_in = slot < P.n_q;
    const float nq = q_in ? P.q_nq[slot] : 0.f;
    const bool q_ok = nq > 0.f;
    const float base = q_ok ? P.q_dotS[slot] + P.q_dotX[slot] : 0.f;
    const float corrS = q_ok ? P.q_corrS[slot] : 0.f;
    const uint32_t *rt_keys = (const uint32_t *)S.rtab;
    const float *rt_w = (const float *)(rt_keys + RT_SLOTS);
    const uint32_t *rt_q = (const uint32_t *)(rt_w + RT_SLOTS);
    const uint32_t *rt_multi = rt_q + RT_SLOTS;
    const int list = (tile * 4 + qtr) * P.n_bsplits + bsplit;
    int *my_pages = S.pages + qtr * P.max_pages;
    float sm[B_SEEDS];
    int sc[B_SEEDS];
#pragma unroll
    for (int i = 0; i < B_SEEDS; i++) { sm[i] = -1.f; sc[i] = -1; }
    unsigned int n_pairs = 0, n_recs = 0;
    int64_t it = 0;
    for (int64_t bk = blk_lo; bk < blk_hi; bk++, it++) {
      const int as = (int)(it & 1);
      const uint32_t aphase = (uint32_t)((it >> 1) & 1);
      const int64_t c0 = bk * B_BN;
      // ---- join: rare entries of this warp's chunks against the tile's rare-feature table ----
      for (int j = warp; j < B_BN; j += B_WORKERS) {
        const int64_t c = c0 + j;
        if (c >= P.n_chunks) break;
        const BlockInfo bi = P.binfo[c];
        const int nr = bi.n_rare;
        const uint32_t *words = P.blk + (size_t)bi.off4 * 4;
        float *Rj = &S.R[j][0];
        uint32_t nxt = lane < nr ? __ldg(words + lane) : PAD_WORD;
        for (int e0 = 0; e0 < nr; e0 += 32) {
          const uint32_t w = nxt;
          const int en = e0 + 32 + lane;
          nxt = en < nr ? __ldg(words + en) : PAD_WORD;
          const uint32_t fid = (w >> 5) & FID_MASK;
          if (fid == FID_NONE) continue;
          uint32_t h = hash_fid(fid, 11);
          for (;;) {
            const uint32_t key = rt_keys[h];
            if (key == KEY_EMPTY) break;
            if (key == fid) {
              uint32_t tf = w & 31u;
              if (tf == TF_OVF) tf = ovf_lookup(P.ovf_keys, P.ovf_vals, P.n_ovf, c, (uint32_t)(e0 + lane));
              const float x = __fmul_ru(rt_w[h], (float)tf);
              const uint32_t qinfo = rt_q[h];
              if (qinfo < (uint32_t)TILE_Q) {
                atomicAdd(&Rj[qinfo], x);
              } else if (qinfo != 0xFFFFFFFFu) {
                const uint32_t *mm = rt_multi + 4 * (qinfo & 0x7FFFFFFFu);
#pragma unroll
                for (int g = 0; g < 4; g++)
                  for (uint32_t bits = mm[g]; bits; bits &= bits - 1) atomicAdd(&Rj[g * 32 + __ffs(bits) - 1], x);
              }
              break;
            }
            h = (h + 1) & (RT_SLOTS - 1);
          }
        }
      }
      if (threadIdx.x < B_BN) S.minB[as][threadIdx.x] = P.chunk_minB[c0 + threadIdx.x];
      // this block's threshold of the query (pass 1: theta0 from the seed scan, possibly raised by peers meanwhile)
      float tq = 0.f;
      if (P.pass == 1 && q_ok) {
        const float th = __int_as_float(__ldcg(&P.gthr[slot]));
        if (th > 0.f) tq = P.jaccard ? th / PRUNE_SLACK : th * th * nq / (PRUNE_SLACK * PRUNE_SLACK);
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");
      // ---- epilogue: frequent part from TMEM + rare part from R -> bound -> seed / candidate ----
      mbar_wait(&S.tmem_full[as], aphase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t v[32];
      {
        const uint32_t taddr = tmem_base + ((uint32_t)(qtr * 32) << 16) + (uint32_t)(as * B_BN + cs * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.tmem_empty[as]);  // the values are in registers: the accumulator may be overwritten
      const float *mb = &S.minB[as][cs * 32];
      float *Rcol = &S.R[cs * 32][qi];
      const int64_t cbase = c0 + cs * 32;
      uint32_t mymask = 0;
#pragma unroll
      for (int j = 0; j < 32; j++) {
        const float x = base + __uint_as_float(v[j]) + Rcol[j * TILE_Q];
        Rcol[j * TILE_Q] = 0.f;
        const bool c_ok = cbase + j < P.n_chunks;
        const float den = P.jaccard ? (nq + mb[j] - x) : (mb[j] + corrS);
        if (P.pass == 1) {
          const float lhs = P.jaccard ? x : x * x;
          const bool sv = q_ok && c_ok && (tq <= 0.f || den <= 0.f || lhs >= tq * den);
          const uint32_t m = __ballot_sync(FULL, sv);
          if (lane == j) mymask = m;
        } else if (q_ok && c_ok) {
          const float lhs = P.jaccard ? x : x * x;
          float metric = den > 0.f ? __fdividef(lhs, den) : INFINITY;
          if (metric > sm[B_SEEDS - 1]) {
            int cc = (int)(cbase + j);
#pragma unroll
            for (int i = 0; i < B_SEEDS; i++)
              if (metric > sm[i]) {
                const float tm = sm[i]; sm[i] = metric; metric = tm;
                const int tc = sc[i]; sc[i] = cc; cc = tc;
              }
          }
        }
      }
      if (P.pass == 1) {
        // lanes holding a non-empty mask append {chunk, mask} to the group's list (warp-aggregated; four warps share
        // a list; the warp whose range crosses into a new page allocates it)
        const uint32_t am = __ballot_sync(FULL, mymask != 0);
        if (am) {
          const int n = __popc(am);
          unsigned int bpos = 0;
          if (lane == 0) {
            bpos = atomicAdd(&S.lcount[qtr], (unsigned int)n);
            for (unsigned int pg = (bpos + PAGE_RECS - 1) / PAGE_RECS; pg * PAGE_RECS < bpos + n; pg++) {
              unsigned int np = atomicAdd(P.pool_next, 1u);
              if (np >= P.pool_pages) { *P.overflow = 1; np = 0; }
              P.list_pages[(size_t)list * P.max_pages + pg] = np;
              __threadfence_block();
              *(volatile int *)&my_pages[pg] = (int)np;
            }
          }
          bpos = __shfl_sync(FULL, bpos, 0);
          if (mymask) {
            const unsigned int pos = bpos + (unsigned int)__popc(am & lanemask_lt());
            const unsigned int pg = pos / PAGE_RECS;
            int page;
            while ((page = *(volatile int *)&my_pages[pg]) < 0) {}
            uint2 rec;
            rec.x = (uint32_t)(cbase + lane);
            rec.y = mymask;
            P.pool[(size_t)page * PAGE_RECS + (pos % PAGE_RECS)] = rec;
            n_pairs += (unsigned int)__popc(mymask);
            n_recs++;
          }
        }
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");  // R is clean again before the next block's join
    }
    if (P.pass == 0) {
      if (q_in) {
        int *o = P.seeds + ((size_t)slot * P.n_bsplits + bsplit) * B_SEEDS_PER_QUERY + cs * B_SEEDS;
#pragma unroll
        for (int i = 0; i < B_SEEDS; i++) o[i] = sc[i];
      }
    } else {
      if (cs == 0 && lane == 0) P.list_count[list] = S.lcount[qtr];
      if (P.stats) {
        for (int o = 16; o; o >>= 1) {
          n_pairs += __shfl_xor_sync(FULL, n_pairs, o);
          n_recs += __shfl_xor_sync(FULL, n_recs, o);
        }
        if (lane == 0) {
          atomicAdd(&P.stats[2], (unsigned long long)n_pairs);
          atomicAdd(&P.stats[3], (unsigned long long)n_recs);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == B_WORKERS + 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
  }
}

typedef CUresult (*PFN_encodeTiled_kv)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                       const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// tensor map of a row-major fp16 matrix [rows][NF], box = 64 columns x 128 rows, 128-byte swizzle
static int make_map_f16_nf(CUtensorMap *map, const void *base, int64_t rows) {
  static PFN_encodeTiled_kv fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p)
      return kv_fail(KV_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    fn = (PFN_encodeTiled_kv)p;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)NF, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)NF * 2};
  cuuint32_t box[2] = {(cuuint32_t)B_BK, 128u};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return kv_fail(KV_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return KV_OK;
}

}  // namespace kvk
