// K1b-B: upper bounds of every (query, chunk) score on the 5th-gen tensor cores, and the candidate lists they leave.
//
// For a query q and a chunk c (32 rows), with U_c(t) = largest tf of feature t in the chunk (0 if absent) and
// Bmin_c = smallest positive row norm of the chunk:
//     dot(q, r) <= dotS_q + dotX_q + sum_{t frequent} W_q(t) U_c(t) + sum_{t rare, in the chunk} w_tile(t) U_c(t)
//     B_r + corr(q, r) >= Bmin_c + corrS_q                                  (corrS_q: every d(t) of q at its largest tf)
// so  score(q, r) <= ub(q, c) = dot_bound / sqrt(|q|^2 (Bmin_c + corrS_q))  for every row r of c  (exact block-max
// pruning: a chunk whose bound is below a query's k-th best score cannot hold a top-k row for it).
//   * the sum over the NF = 256 FREQUENT features is a [128 queries x 256] x [256 x 64 chunks] fp16 GEMM (weights
//     rounded UP to fp16, tf exact): tcgen05.mma cta_group::1 kind::f16, M = 128, N = 64, fp32 accumulators in TMEM
//     (double buffered), the query operand resident in shared memory for the CTA's life, the chunk operand streamed
//     by TMA (128B swizzle, mbarrier expect_tx) -- this only computes BOUNDS; scores stay exact integer sums (K1b-S);
//   * the next NF2 = 1024 features by chunk frequency ("second class": mid-frequency words and bigrams, each shared by
//     many queries of a tile) are kept per 64-chunk block as two TRANSPOSED presence bitmaps Ubt[feature][tf >= 1,
//     tf >= 2][64 chunk bits] (16 KB, one bulk-async copy per block, double buffered): a query thread reads ONE word per
//     feature of its own list (<= Q2CAP) and has that feature's presence in its 16 chunk columns -- no atomics;
//   * the remaining RARE features are joined the other way round: every 64-chunk block carries (built at finalize) a
//     presence bitmap and an open-addressing table of its rare features (feature -> mask of the block's chunks holding
//     it, largest tf); each query looks ITS OWN <= 32 rare features up -- one shared-memory bit test per (query,
//     feature, block), and only on a hit a probe of the block's table in L2 -- and adds weight x tf into
//     R[chunk][query] (shared memory).  1.8k bit tests per tile and block instead of 5.4k entry probes;
//   * epilogue (16 warps, four threads per query, tcgen05.ld of 16 chunk columns each): bound vs the query's threshold.
//     pass 0 keeps, per thread, the 4 best chunks by bound (16 seeds per query: K1b-S scores them first, which gives
//     every query a close lower bound theta0 of its k-th best score) and stores every bound as an 8-bit code rounded up
//     (the selection kernel below builds the candidate lists from the codes); pass 1 (only when the codes do not fit in
//     memory) recomputes the bounds and appends {chunk, mask of the group's surviving queries} to the scan group's
//     candidate list (paged pool) for every chunk with ub >= theta0.
// One CTA = one 128-query tile x one range of 64-chunk blocks; 18 warps: 16 workers (join + epilogue), TMA, MMA.
#pragma once
#include "tfidf_kernels.cuh"

#include <cuda.h>

namespace kvk {

constexpr int B_BN = 64;                     // chunks per block = N of the MMA tile
constexpr int B_BK = 64;                     // K slice: 64 fp16 = one 128-byte swizzle row
constexpr int B_KSLICES = NF / B_BK;         // 4
constexpr int B_STAGES = 2;
constexpr int B_A_SLICE_BYTES = TILE_Q * B_BK * 2;  // 16 KiB: one K slice of the query operand
constexpr int B_B_SLICE_BYTES = B_BN * B_BK * 2;    // 8 KiB: one K slice of the chunk operand
constexpr int B_WORKERS = 16;
constexpr int B_THREADS = (B_WORKERS + 2) * 32;
constexpr int B_COLS = B_BN / 4;             // chunk columns per epilogue thread (four threads serve a query)
constexpr int B_SEEDS = 4;                   // seeds per worker thread
constexpr int B_SEEDS_PER_QUERY = 4 * B_SEEDS;
constexpr float UBQ_SCALE = 250.f;           // 8-bit bound code = ceil(bound * 250), saturating at 255 (bounds reach ~1.001)

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

// instruction descriptor: D = f32, A = B = f16, both K-major, N = 64, M = 128
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
// wait of a single-lane role (TMA producer, MMA issuer): polls with a pause, so that the spinning lane does not
// take issue slots from the worker warps sharing its scheduler
__device__ __forceinline__ void mbar_wait_idle(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  for (;;) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(64);
  }
}

// MUFU.RSQ without the denormal pre-/post-scaling rsqrtf() carries (the argument is |q|^2 x a row norm: far from denormal;
// for normal arguments the result is the same bit pattern)
__device__ __forceinline__ float rsqrt_approx(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
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
  const uint2 *q3list;                                    // [n_tiles][Q3CAP][TILE_Q] rare (feature, weight) lists of the queries
  const uint32_t *rbloom;                                 // [n_blocks][RB_BITS / 32]
  const uint32_t *rt_keys;                                // per-block rare tables: feature << 5 | largest tf (31: tfmax[])
  const unsigned long long *rt_masks;                     // ... chunks of the block holding the feature
  const uint32_t *rt_off, *rt_size;                       // [n_blocks]
  const uint32_t *tfmax;                                  // [V] largest tf of a feature (for entries whose 5-bit tf overflowed)
  const uint2 *q2list;                                    // [n_tiles][Q2CAP][TILE_Q]
  const uint32_t *ubt;                                    // [n_blocks][NF2][2][B_BN / 32]
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
  unsigned char *ubq;         // pass 0, optional: every bound as an 8-bit code (rounded up), [n_q][ubq_stride]
  int64_t ubq_stride;
  float *dbg_xs;              // test hook: when set, the numerator of every bound, [n_q][n_chunks_pad]
  int64_t dbg_stride;
};

struct __align__(1024) BoundSmem {
  unsigned char a[B_KSLICES][B_A_SLICE_BYTES];   // the tile's weight rows, resident
  unsigned char b[B_STAGES][B_B_SLICE_BYTES];    // chunk slices in flight
  uint32_t ubt[2][NF2][2][B_BN / 32];            // second-class bitmaps (tf >= 1, tf >= 2) of the current / next block of chunks
  uint32_t rbm[2][RB_BITS / 32];                 // rare-feature presence bitmap of the current / next block
  float R[B_BN][TILE_Q];                         // rare part of the dot bound, [chunk][query]
  uint2 q2[Q2CAP][TILE_Q];                       // the queries' second-class lists (bit row | (tfmax - 1) << 16, weight)
  uint2 q3[Q3CAP][TILE_Q];                       // the queries' rare lists (feature id, weight)
  float minB[2][B_BN];
  uint64_t full_bar[B_STAGES], empty_bar[B_STAGES], a_bar, blk_bar[2], tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
  unsigned int lcount[4];
  int pages[1];  // [4][max_pages], sized at launch
};

static inline size_t bound_smem_bytes(int max_pages) { return sizeof(BoundSmem) + (size_t)4 * max_pages * sizeof(int) + 1024; }

// FAST = the configuration of the headline path fixed at compile time (pass 0 with bound codes, TF-IDF cosine, no test
// hook): the epilogue then holds no uniform branches, parameter reloads or dead variants.  FAST = false is the same code
// with those four switches read from the parameters.
template <bool FAST>
__global__ void __launch_bounds__(B_THREADS, 1)
tfidf_bound_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_u, BoundParams P) {
  extern __shared__ unsigned char smem_raw[];
  const int pass = FAST ? 0 : P.pass;
  const bool jacc = FAST ? false : (P.jaccard != 0);
  const bool has_codes = FAST ? true : (P.ubq != nullptr);
  // 1024-byte alignment by an OFFSET into the shared array (not by rounding a generic pointer): the compiler keeps
  // the shared address space, so every access below is LDS/STS/ATOMS instead of a generic load / store / atomic
  BoundSmem &S = *reinterpret_cast<BoundSmem *>(smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, bsplit = blockIdx.y;
  const int64_t n_blocks = (P.n_chunks + B_BN - 1) / B_BN;
  const int64_t blk_lo = n_blocks * bsplit / P.n_bsplits, blk_hi = n_blocks * (bsplit + 1) / P.n_bsplits;

  if (threadIdx.x == 0) {
    for (int i = 0; i < B_STAGES; i++) { mbar_init(&S.full_bar[i], 1); mbar_init(&S.empty_bar[i], 1); }
    mbar_init(&S.a_bar, 1);
    mbar_init(&S.blk_bar[0], 1);
    mbar_init(&S.blk_bar[1], 1);
    for (int i = 0; i < 2; i++) { mbar_init(&S.tmem_full[i], 1); mbar_init(&S.tmem_empty[i], B_WORKERS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == B_WORKERS + 1) {  // TMEM: 128 columns = two 128x64 fp32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_addr(&S.tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {  // second-class and rare lists of the tile's queries, R = 0, list state
    const uint4 *src2 = (const uint4 *)(P.q2list + (size_t)tile * Q2CAP * TILE_Q);
    uint4 *dst2 = (uint4 *)&S.q2[0][0];
    for (int i = threadIdx.x; i < Q2CAP * TILE_Q / 2; i += B_THREADS) dst2[i] = src2[i];
    const uint4 *src3 = (const uint4 *)(P.q3list + (size_t)tile * Q3CAP * TILE_Q);
    uint4 *dst3 = (uint4 *)&S.q3[0][0];
    for (int i = threadIdx.x; i < Q3CAP * TILE_Q / 2; i += B_THREADS) dst3[i] = src3[i];
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
      mbar_expect_tx(&S.a_bar, B_KSLICES * B_A_SLICE_BYTES);
      for (int s = 0; s < B_KSLICES; s++) tma_load_2d(S.a[s], &map_w, &S.a_bar, s * B_BK, tile * TILE_Q);
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t bk = blk_lo; bk < blk_hi; bk++) {
        for (int s = 0; s < B_KSLICES; s++) {
          mbar_wait_idle(&S.empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&S.full_bar[stage], B_B_SLICE_BYTES);
          tma_load_2d(S.b[stage], &map_u, &S.full_bar[stage], s * B_BK, (int)(bk * B_BN));
          if (++stage == B_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == B_WORKERS + 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      mbar_wait_idle(&S.a_bar, 0);
      int stage = 0;
      uint32_t phase = 0;
      int64_t it = 0;
      for (int64_t bk = blk_lo; bk < blk_hi; bk++, it++) {
        const int as = (int)(it & 1);
        const uint32_t aphase = (uint32_t)((it >> 1) & 1);
        mbar_wait_idle(&S.tmem_empty[as], aphase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * B_BN);
        for (int s = 0; s < B_KSLICES; s++) {
          mbar_wait_idle(&S.full_bar[stage], phase);
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
    const int list = (tile * 4 + qtr) * P.n_bsplits + bsplit;
    int *my_pages = S.pages + qtr * P.max_pages;
    float sm[B_SEEDS];
    int sc[B_SEEDS];
#pragma unroll
    for (int i = 0; i < B_SEEDS; i++) { sm[i] = -1.f; sc[i] = -1; }
    unsigned int n_pairs = 0, n_recs = 0;
    // per-block side data (second-class bitmaps + rare presence bitmap): bulk-async copies, double buffered -- block
    // b + 1 is fetched while block b is processed
    auto fetch_block = [&](int64_t bk2, int buf) {
      mbar_expect_tx(&S.blk_bar[buf], (uint32_t)(sizeof(S.ubt[0]) + sizeof(S.rbm[0])));
      bulk_copy_g2s(&S.ubt[buf][0][0][0], P.ubt + (size_t)bk2 * (sizeof(S.ubt[0]) / 4), sizeof(S.ubt[0]), &S.blk_bar[buf]);
      bulk_copy_g2s(&S.rbm[buf][0], P.rbloom + (size_t)bk2 * (RB_BITS / 32), sizeof(S.rbm[0]), &S.blk_bar[buf]);
    };
    if (threadIdx.x == 0 && blk_lo < blk_hi) fetch_block(blk_lo, 0);
    int64_t it = 0;
    for (int64_t bk = blk_lo; bk < blk_hi; bk++, it++) {
      const int as = (int)(it & 1);
      const uint32_t aphase = (uint32_t)((it >> 1) & 1);
      const int64_t c0 = bk * B_BN;
      if (threadIdx.x == 0 && bk + 1 < blk_hi) fetch_block(bk + 1, as ^ 1);  // its buffers were last read two barriers ago
      mbar_wait(&S.blk_bar[as], aphase);
      // ---- rare join, inverted: thread (query, quarter of its rare list) tests each feature in the block's presence
      //      bitmap; on a hit it probes the block's table (L2) and adds weight x tf to R[chunk][query] for every chunk
      //      of the mask.  Rows are text-sorted, so one feature can sit in most chunks of a block: R is dense. ----
      {
        const int jq = threadIdx.x & (TILE_Q - 1), part = threadIdx.x >> 7;  // 512 worker threads = 128 queries x 4
        const uint32_t *bm = S.rbm[as];
        const uint32_t tsize = P.rt_size[bk], toff = P.rt_off[bk];
#pragma unroll 1
        for (int i = part; i < Q3CAP; i += 4) {
          const uint2 f3 = S.q3[i][jq];
          if (f3.x >= FID_NONE) break;  // the lists are filled from the front (padding queries: all ones)
          const uint32_t bb = rb_bit(f3.x);
          if (!((bm[bb >> 5] >> (bb & 31u)) & 1u)) continue;
          uint32_t h = rt_slot(f3.x, tsize);
          for (;;) {
            const uint32_t key = __ldg(P.rt_keys + toff + h);
            if (key == KEY_EMPTY) break;
            if ((key >> 5) == f3.x) {
              uint32_t tf = key & 31u;
              if (tf == TF_OVF) tf = __ldg(P.tfmax + f3.x);
              const float x = __fmul_ru(__uint_as_float(f3.y), (float)tf);
              unsigned long long cm = __ldg(P.rt_masks + toff + h);
              while (cm) {
                const int j = __ffsll((long long)cm) - 1;
                cm &= cm - 1;
                atomicAdd(&S.R[j][jq], x);
              }
              break;
            }
            h = h + 1 == tsize ? 0 : h + 1;
          }
        }
      }
      if (threadIdx.x < B_BN) S.minB[as][threadIdx.x] = P.chunk_minB[c0 + threadIdx.x];
      // this block's threshold of the query (pass 1: theta0 from the seed scan, possibly raised by peers meanwhile)
      float tq = 0.f;
      if (pass == 1 && q_ok) {
        const float th = __int_as_float(__ldcg(&P.gthr[slot]));
        if (th > 0.f) tq = jacc ? th / PRUNE_SLACK : th * th * nq / (PRUNE_SLACK * PRUNE_SLACK);
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");
      // ---- epilogue, thread = query, B_COLS chunk columns: rare part (R) + second-class part (bitmaps) + frequent
      //      part (TMEM) -> bound -> seed / candidate ----
      float x[B_COLS];
      {
        float *Rcol = &S.R[cs * B_COLS][qi];
#pragma unroll
        for (int j = 0; j < B_COLS; j++) { x[j] = Rcol[j * TILE_Q]; Rcol[j * TILE_Q] = 0.f; }
      }
      const uint32_t bsh = (uint32_t)((cs * B_COLS) & 31), bword = (uint32_t)((cs * B_COLS) >> 5);
      constexpr uint32_t CMASK = B_COLS == 32 ? 0xFFFFFFFFu : ((1u << B_COLS) - 1u);
#pragma unroll 1
      for (int i = 0; i < Q2CAP; i++) {
        const uint2 f2 = S.q2[i][qi];
        const float w2 = q_ok ? __uint_as_float(f2.y) : 0.f;
        if (!__any_sync(FULL, w2 > 0.f)) break;  // the lists are filled from the front
        const uint32_t row2 = f2.x & 0xFFFFu, tm1 = f2.x >> 16;
        const uint32_t m = w2 > 0.f ? ((S.ubt[as][row2][0][bword] >> bsh) & CMASK) : 0u;
        if (m) {
#pragma unroll
          for (int j = 0; j < B_COLS; j++)
            if ((m >> j) & 1u) x[j] += w2;
          const uint32_t mm = tm1 ? ((S.ubt[as][row2][1][bword] >> bsh) & CMASK) : 0u;  // tf >= 2 there: up to tfmax - 1 more
          if (mm) {
            const float wex = __fmul_ru(w2, (float)tm1);
#pragma unroll
            for (int j = 0; j < B_COLS; j++)
              if ((mm >> j) & 1u) x[j] += wex;
          }
        }
      }
      mbar_wait(&S.tmem_full[as], aphase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t v[B_COLS];
      {
        static_assert(B_COLS == 16, "the TMEM load below reads 16 columns");
        const uint32_t taddr = tmem_base + ((uint32_t)(qtr * 32) << 16) + (uint32_t)(as * B_BN + cs * B_COLS);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.tmem_empty[as]);  // the values are in registers: the accumulator may be overwritten
      const float *mb = &S.minB[as][cs * B_COLS];
      const int64_t cbase = c0 + cs * B_COLS;
      uint32_t mymask = 0;
      const bool dbg = FAST ? false : (P.dbg_xs != nullptr);
      const int nv = (int)min((int64_t)B_COLS, P.n_chunks - cbase);  // chunk columns of this thread that exist
      uint32_t codes[B_COLS / 4];
#pragma unroll
      for (int j = 0; j < B_COLS / 4; j++) codes[j] = 0;
#pragma unroll
      for (int j = 0; j < B_COLS; j++) {
        const float xs = base + __uint_as_float(v[j]) + x[j];
        const bool c_ok = j < nv;
        if (dbg && q_in && c_ok) P.dbg_xs[(size_t)slot * P.dbg_stride + cbase + j] = xs;
        const float den = jacc ? (nq + mb[j] - xs) : (mb[j] + corrS);
        if (pass == 1) {
          const float lhs = jacc ? xs : xs * xs;
          const bool sv = q_ok && c_ok && (tq <= 0.f || den <= 0.f || lhs >= tq * den);
          const uint32_t m = __ballot_sync(FULL, sv);
          if (lane == j) mymask = m;
        } else if (q_ok && c_ok) {
          // the bound itself (slack included): seeds are ranked by it, and it is stored as a code that only errs upwards
          float metric = den > 0.f ? (jacc ? __fdividef(xs, den) : xs * rsqrt_approx(nq * den)) * (PRUNE_SLACK * 1.00001f) : INFINITY;
          if (has_codes) codes[j >> 2] |= (uint32_t)fminf(255.f, ceilf(metric * UBQ_SCALE)) << ((j & 3) * 8);
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
      if (pass == 0 && has_codes && q_in)
        *reinterpret_cast<uint4 *>(P.ubq + (size_t)slot * P.ubq_stride + cbase) = make_uint4(codes[0], codes[1], codes[2], codes[3]);
      if (pass == 1) {
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
      asm volatile("bar.sync 1, 512;" ::: "memory");  // R is clean again and every reader of this block's side data is done
    }
    if (pass == 0) {
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_base) : "memory");
  }
}

// ----------------------------------------------------------------------------------------
// K1b-B2: candidate lists from the stored 8-bit bound codes (second pass without recomputing the bounds).  One CTA =
// one scan group (32 queries = the lanes) x one chunk range; a warp reads, per query, 32 codes (one 32-byte sector) and
// compares them with the query's threshold code; ballots give the group's query mask per chunk; non-empty masks are
// appended to the group's paged candidate list exactly as K1b-B's pass 1 does.  HBM-bound: n_q x chunks bytes read once.
// ----------------------------------------------------------------------------------------
struct SelectParams {
  const unsigned char *ubq;
  int64_t ubq_stride, n_chunks, n_q;
  const float *q_nq;
  const int *gthr;
  int n_bsplits;
  uint32_t *list_count;
  uint32_t *list_pages;
  int max_pages;
  uint2 *pool;
  unsigned int *pool_next;
  unsigned int pool_pages;
  int *overflow;
  unsigned long long *stats;
};

constexpr int SEL_WARPS = 8;

__global__ void __launch_bounds__(SEL_WARPS * 32) tfidf_select_kernel(SelectParams P) {
  extern __shared__ int s_pages[];  // [max_pages]
  __shared__ unsigned int s_count;
  const int group = blockIdx.x, bsplit = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int list = group * P.n_bsplits + bsplit;
  for (int i = threadIdx.x; i < P.max_pages; i += blockDim.x) s_pages[i] = -1;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  const int64_t slot = (int64_t)group * GROUP_Q + lane;
  const bool q_ok = slot < P.n_q && P.q_nq[slot] > 0.f;
  uint32_t tcode = 256;  // never reached: the query takes no candidates
  if (q_ok) {
    const float th = __int_as_float(P.gthr[slot]);
    tcode = th > 0.f ? (uint32_t)fminf(255.f, floorf(th * UBQ_SCALE)) : 0u;
  }
  // the chunk range of this split, in units of 32 chunks, interleaved over the warps
  const int64_t n_blocks = (P.n_chunks + B_BN - 1) / B_BN;
  const int64_t c_lo = (n_blocks * bsplit / P.n_bsplits) * B_BN, c_hi = min(P.n_chunks, (n_blocks * (bsplit + 1) / P.n_bsplits) * B_BN);
  const unsigned char *row = P.ubq + (size_t)min(slot, P.n_q - 1) * P.ubq_stride;
  unsigned int n_pairs = 0, n_recs = 0;
  for (int64_t c = c_lo + 32 * warp; c < c_hi; c += 32 * SEL_WARPS) {
    const uint4 a = __ldcs(reinterpret_cast<const uint4 *>(row + c)), b = __ldcs(reinterpret_cast<const uint4 *>(row + c + 16));
    const uint32_t wds[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t mymask = 0;
#pragma unroll
    for (int j = 0; j < 32; j++) {
      const uint32_t code = (wds[j >> 2] >> ((j & 3) * 8)) & 0xFFu;
      const uint32_t m = __ballot_sync(FULL, q_ok && code >= tcode && c + j < c_hi);
      if (lane == j) mymask = m;
    }
    const uint32_t am = __ballot_sync(FULL, mymask != 0);
    if (am) {
      const int n = __popc(am);
      unsigned int bpos = 0;
      if (lane == 0) {
        bpos = atomicAdd(&s_count, (unsigned int)n);
        for (unsigned int pg = (bpos + PAGE_RECS - 1) / PAGE_RECS; pg * PAGE_RECS < bpos + n; pg++) {
          unsigned int np = atomicAdd(P.pool_next, 1u);
          if (np >= P.pool_pages) { *P.overflow = 1; np = 0; }
          P.list_pages[(size_t)list * P.max_pages + pg] = np;
          __threadfence_block();
          *(volatile int *)&s_pages[pg] = (int)np;
        }
      }
      bpos = __shfl_sync(FULL, bpos, 0);
      if (mymask) {
        const unsigned int pos = bpos + (unsigned int)__popc(am & lanemask_lt());
        const unsigned int pg = pos / PAGE_RECS;
        int page;
        while ((page = *(volatile int *)&s_pages[pg]) < 0) {}
        uint2 rec;
        rec.x = (uint32_t)(c + lane);
        rec.y = mymask;
        P.pool[(size_t)page * PAGE_RECS + (pos % PAGE_RECS)] = rec;
        n_pairs += (unsigned int)__popc(mymask);
        n_recs++;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) P.list_count[list] = s_count;
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

typedef CUresult (*PFN_encodeTiled_kv)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                       const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// tensor map of a row-major fp16 matrix [rows][NF], box = 64 columns x box_rows rows, 128-byte swizzle
static int make_map_f16_nf(CUtensorMap *map, const void *base, int64_t rows, int box_rows) {
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
  cuuint32_t box[2] = {(cuuint32_t)B_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return kv_fail(KV_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return KV_OK;
}

}  // namespace kvk
