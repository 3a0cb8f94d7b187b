// K1 (main pass): fused InfoNCE logits + softmax statistics + d loss/d q on the
// 5th-generation tensor cores (tcgen05, tf32 operands, fp32 accumulate in TMEM),
// queue tiles staged by TMA.  sm_100a only.
//
// Mathematically this is one attention forward with V == K:
//      S = Q W^T (scaled per key by dscale_j), P = 2^(S - shift), O = P (dscale . W)
// so the row sums of P give the log-sum-exp and O gives sum_j p_ij decay_j queue_j,
// the negative part of d loss/d q, in a SINGLE pass over the queue (the reference
// makes 3 passes for the decayed snapshot plus a GEMM, and autograd a second GEMM).
//
// CTA = 192 threads, one CTA per SM, 128 query rows x a contiguous range of 64-key tiles:
//   warp 0      TMA producer   queue tile [64 keys x 128 ch] fp32 -> smem as 4 channel blocks of
//                              [64][32 ch = 128 B], TWICE: once 128B-swizzled (K-major operand of
//                              MMA1) and once with the 32-byte-atom 128B swizzle, the only layout
//                              tcgen05 accepts for an MN-major tf32 operand (MMA2); the second
//                              read hits L2.  Plus the tile's 64 dscale floats (bulk copy).
//   warp 1      MMA issuer     MMA1: S[128 x 64]  = Q[128 x 128] . Wt     (TS: A = Q in TMEM, B K-major)
//                              MMA2: O[128 x 128] += P[128 x 64] . W      (TS: A = P in TMEM, B MN-major)
//   warps 2..5  softmax        thread <-> query row (TMEM lane): stage q into TMEM once, then per
//                              tile tcgen05.ld S, exp2, row sum, count(s > pos), P' = p * dscale
//                              -> tcgen05.st over S
// TMEM (512 columns): O = [0,128), S/P double buffer = [128,192) [192,256), Q = [256,384).
// Pipelines: full/empty (TMA <-> MMA, 3 stages), q_full, s_full (MMA1 -> softmax),
// p_full (softmax -> MMA2), o_full (last MMA2 -> epilogue).
#include "tc_common.cuh"

namespace mscl {
namespace tc {

constexpr int kTile = 128;            // keys per MMA1 dispatch / softmax step (a tcgen05 dispatch costs >= ~60 cycles
                                      // whatever N is, so N = 128 keys halves the MMA1 time of N = 64)
constexpr int kHalf = 64;             // keys per ring-2 slot (MMA2 consumes a tile in two halves)
constexpr int kStages1 = 2;           // ring 1: K-major copies, 128 keys each
constexpr int kStages2 = 3;           // ring 2: MN-major copies, 64 keys each
constexpr int kCb = 4;                // channel blocks of 32 fp32 (one 128-byte swizzle row)
constexpr int kSoftmaxWarps = 8;      // two per TMEM lane quarter, each takes 64 of a tile's 128 keys
constexpr int kThreads = (3 + kSoftmaxWarps) * 32;   // ring-1 producer, MMA issuer, softmax warps, ring-2 producer

constexpr uint32_t kW1Bytes = kTile * kC * 4;         // 65536: K-major copy of a 128-key tile
constexpr uint32_t kW1Slab = kTile * 128;             // bytes per channel block of it
constexpr uint32_t kW2Bytes = kHalf * kC * 4;         // 32768: MN-major copy of a 64-key half tile
constexpr uint32_t kW2Slab = kHalf * 128;
constexpr uint32_t kQBytes = kRows * kC * 4;          // 65536: the Q tile (and the O tile on the way out)
constexpr uint32_t kQSlab = kRows * 128;              // bytes per channel block of the Q tile

// shared memory map (offsets from the 1024-aligned base)
// Two rings with separate lifetimes: ring 1 holds the 128B-swizzled (K-major) copy of a tile, needed by
// MMA1(t) only; ring 2 the 32B-atom-swizzled (MN-major) copy, needed by MMA2(t) one tile period later.
// A ring-1 slot is free as soon as MMA1 has read it, so the HBM prefetch runs ahead of MMA1; ring 2 is
// filled from L2 (the same bytes again).
constexpr uint32_t kOffW = 0;                             // ring 1
constexpr uint32_t kOffW2 = kOffW + kStages1 * kW1Bytes;  // ring 2
constexpr uint32_t kOffBar = kOffW2 + kStages2 * kW2Bytes;
constexpr uint32_t kNumBars = 2 * kStages1 + 2 * kStages2 + 1 + 2 + 2 + 1 + 2;  // full1, empty1, full2, empty2, q, s_full[2], p_full[2], o_full, qload, qfree
constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr uint32_t kOffRed = kOffTmemPtr + 16;          // [2][128] floats: sum / count of the second column half
constexpr uint32_t kSmemUsed = kOffRed + 2 * kRows * 4;
static_assert(kQBytes <= 2 * kW2Bytes && kStages2 >= 3, "the Q tile is staged through ring-2 slots 1 and 2");
static_assert(kQBytes <= kW1Bytes, "the O tile is staged through ring-1 slot 0");
constexpr uint32_t kSmemBytes = kSmemUsed + 1024;     // slack for manual 1024-byte alignment
static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory a CTA may use");

constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColO = 0;
constexpr uint32_t kColS = 128;       // two S/P buffers of kTile columns
constexpr uint32_t kColQ = 384;

// instruction descriptors (cute::UMMA::InstrDescriptor bit layout):
//  [4,6) c_format=1 (f32) | [7,10) a_format=2 (tf32) | [10,13) b_format=2 (tf32)
//  [15] a_major | [16] b_major (1 = MN-major) | [17,23) N>>3 | [24,29) M>>4
constexpr uint32_t kIdescBase = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kRows >> 4) << 24);
constexpr uint32_t kIdesc1 = kIdescBase | ((uint32_t)(kTile >> 3) << 17);               // N = 128 keys
constexpr uint32_t kIdesc2 = kIdescBase | (1u << 16) | ((uint32_t)(kC >> 3) << 17);     // N = 128, B MN-major


#ifdef MSCL_TC_TIMELINE
// debug build only (MSCL_TIMELINE=1 python -m mscl_b200.build): per-CTA phase timestamps
__device__ unsigned long long g_timeline[148 * 2 * 32];
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TL(slot) g_timeline[(blockIdx.y * gridDim.x + blockIdx.x) * 32 + (slot)] = gtime()
#else
#define TL(slot)
#endif

// 32 keys of one tile for one query row, on registers: S -> p, row sum, hit count, P' (written back over S by the caller).
// `ds` = the 32 per-key scales of this half tile; nvalid / dupcol are relative to the half tile.
// Per element on the fast path: FFMA (logit - shift), MUFU.EX2, FFMA (pos - logit: its sign bit is the
// top-k hit), one add of that sign bit, FADD (row sum), FMUL (p * scale), IADD (round to nearest tf32:
// +half ulp, the tensor core drops the low 13 bits).  4 independent sum / count chains.
template <bool GRAD, bool FULL>
__device__ __forceinline__ void softmax_half(uint32_t (&v)[32], const float4 *ds, float shift2, float pos2,
                                             int nvalid, int dupcol, float &sum, int &cnt) {
  float s4[4] = {0.f, 0.f, 0.f, 0.f};
  uint32_t c4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 d4 = ds[j4];
    const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = j4 * 4 + e;
      const float sv = __uint_as_float(v[j]);
      float p = ex2(fmaf(sv, dd[e], -shift2));
      uint32_t hit = __float_as_uint(fmaf(-sv, dd[e], pos2)) >> 31;    // 1 iff logit > positive logit
      if (!FULL) {  // tail tile (keys beyond K_local) or the tile holding this row's own positive key
        const bool ok = j < nvalid && j != dupcol;      // the duplicate of the positive is added exactly by finalize
        p = ok ? p : 0.f;
        hit = ok ? hit : 0u;
      }
      s4[e] += p;
      c4[e] += hit;
      v[j] = GRAD ? __float_as_uint(p * dd[e]) + 0x1000u : 0u;
    }
  }
  sum += (s4[0] + s4[1]) + (s4[2] + s4[3]);
  cnt += (int)((c4[0] + c4[1]) + (c4[2] + c4[3]));
}

// part: per-CTA partial rows, float [gridDim.x][M][kLd]: O[0:128] | sum-exp | #neg>pos | 0 | 0.
// Every (blockIdx.x, row < M) row is written exactly once (plain stores, no atomics): the cross-CTA sum
// is taken in a fixed order by the finalize / reduce kernels, so results are bit-reproducible.
template <bool GRAD>
__global__ void __launch_bounds__(kThreads, 1)
infonce_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_w2,
                  const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_part,
                  const float *__restrict__ qpack, int M,
                  const float *__restrict__ dscale, int64_t K_local, int64_t shard_begin,
                  float *__restrict__ part) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sW = base + kOffW;
  const uint32_t sW2 = base + kOffW2;
  const uint32_t sQ = sW2 + kW2Bytes;                          // Q tile: ring-2 slots 1 and 2
  const uint32_t bar0 = base + kOffBar;
  auto bar_full1 = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty1 = [&](int s) { return bar0 + 8u * (kStages1 + s); };
  auto bar_full2 = [&](int s) { return bar0 + 8u * (2 * kStages1 + s); };
  auto bar_empty2 = [&](int s) { return bar0 + 8u * (2 * kStages1 + kStages2 + s); };
  constexpr int kB = 2 * kStages1 + 2 * kStages2;
  const uint32_t bar_q = bar0 + 8u * kB;                       // Q rows stored to TMEM (8 warps)
  auto bar_sfull = [&](int b) { return bar0 + 8u * (kB + 1 + b); };
  auto bar_pfull = [&](int b) { return bar0 + 8u * (kB + 3 + b); };
  const uint32_t bar_ofull = bar0 + 8u * (kB + 5);
  const uint32_t bar_qload = bar0 + 8u * (kB + 6);             // Q tile landed in smem (TMA)
  const uint32_t bar_qfree = bar0 + 8u * (kB + 7);             // Q tile read out of smem (8 warps)
  volatile uint32_t *tmem_ptr_smem = reinterpret_cast<volatile uint32_t *>(gbase + kOffTmemPtr);
  float *red_smem = reinterpret_cast<float *>(gbase + kOffRed);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;

  // this CTA's tile range
  const int64_t n_tiles = (K_local + kTile - 1) / kTile;
  const int64_t t_begin = n_tiles * blockIdx.x / gridDim.x;
  const int64_t t_end = n_tiles * (blockIdx.x + 1) / gridDim.x;
  const int nt = (int)(t_end - t_begin);
  const int row0 = blockIdx.y * kRows;

  if (warp == 0 && lane == 0) {
    TL(0);
#ifdef MSCL_TC_TIMELINE
    g_timeline[(blockIdx.y * gridDim.x + blockIdx.x) * 32 + 30] = clock64();
#endif
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w2) : "memory");
    if (GRAD) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_part) : "memory");
    for (int s = 0; s < kStages1; ++s) {
      mbar_init(bar_full1(s), 1);
      mbar_init(bar_empty1(s), 1);
    }
    for (int s = 0; s < kStages2; ++s) {
      mbar_init(bar_full2(s), 1);
      mbar_init(bar_empty2(s), 1);
    }
    mbar_init(bar_q, kSoftmaxWarps);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_sfull(b), 1);
      mbar_init(bar_pfull(b), kSoftmaxWarps);
    }
    mbar_init(bar_ofull, 1);
    mbar_init(bar_qload, 1);
    mbar_init(bar_qfree, kSoftmaxWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     base + kOffTmemPtr),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
  pdl_trigger();   // the finalize kernel may be launched; it waits for this grid's completion before reading the slabs

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      TL(1);
      auto load1 = [&](int t) {      // K-major copy of the 128-key tile t (from HBM)
        const int s = t % kStages1;
        mbar_wait(bar_empty1(s), ((uint32_t)(t / kStages1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(bar_full1(s), kW1Bytes);
        tma_load_3d(sW + s * kW1Bytes, &tmap_w, bar_full1(s), 0, (int)((t_begin + t) * kTile), 0);
      };
      // The queue is older than the prep kernel this grid may overlap with (PDL): start streaming it
      // before waiting for prep's qpack / dscale.
      for (int t = 0; t < kStages1 && t < nt; ++t) load1(t);
      pdl_wait();
      // Q tile [128 rows x 128 ch] as 4 channel-block slabs of [128][128 B], 128B-swizzled; rows >= M are
      // zero-filled by the TMA unit
      mbar_arrive_expect_tx(bar_qload, kQBytes);
      tma_load_3d(sQ, &tmap_q, bar_qload, 0, row0, 0);
      // ring 1 only: a slot is re-armed the moment MMA1 has read it, independent of ring 2's progress
      for (int t = kStages1; t < nt; ++t) load1(t);
      TL(4);
    }
    __syncwarp();
  } else if (warp == 2 + kSoftmaxWarps) {
    // ===================== ring-2 producer (own warp: its waits on MMA2 never hold back the HBM prefetch) ======
    if (GRAD && elect_one()) {
      for (int u = 0; u < 2 * nt; ++u) {   // MN-major copy of the 64-key half tile u (the same bytes again: an L2 hit)
        const int s = u % kStages2;
        if (u == 1) mbar_wait(bar_qfree, 0);        // slots 1 and 2 held the Q tile
        mbar_wait(bar_empty2(s), ((uint32_t)(u / kStages2) & 1u) ^ 1u);
#ifndef MSCL_EXP_NOLOAD2
        mbar_arrive_expect_tx(bar_full2(s), kW2Bytes);
        tma_load_3d(sW2 + s * kW2Bytes, &tmap_w2, bar_full2(s), 0, (int)(t_begin * kTile + (int64_t)u * kHalf), 0);
#else
        mbar_arrive(bar_full2(s));
#endif
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // descriptor words: lo = (addr >> 4) | (LBO >> 4) << 16, hi = (SBO >> 4) | version 1 << 14 | layout << 29
    constexpr uint32_t kHi1 = (1024u >> 4) | (1u << 14) | (2u << 29);      // K-major, SWIZZLE_128B, SBO 1024
    constexpr uint32_t kHi2 = (512u >> 4) | (1u << 14) | (1u << 29);       // MN-major, SWIZZLE_128B_BASE32B, SBO 512
    const uint32_t lo1_base = ((sW & 0x3FFFFu) >> 4) | ((16u >> 4) << 16);              // LBO 16
    const uint32_t lo2_base = ((sW2 & 0x3FFFFu) >> 4) | ((kW2Slab >> 4) << 16);         // LBO = channel-block pitch
    if (elect_one()) {
      auto issue_mma1 = [&](int t) {
        const int s = t % kStages1;
        mbar_wait(bar_full1(s), (uint32_t)(t / kStages1) & 1u);
        tc_fence_after();
        const uint32_t d = tmem + kColS + (uint32_t)(t & 1) * kTile;
        const uint32_t lo = lo1_base + (uint32_t)s * (kW1Bytes >> 4);
#pragma unroll
        for (int cb = 0; cb < kCb; ++cb) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            mma_ts_lh(d, tmem + kColQ + cb * 32 + ks * 8, lo + ((cb * kW1Slab + ks * 32) >> 4), kHi1, kIdesc1,
                      (cb | ks) ? 1u : 0u);
        }
        tc_commit(bar_sfull(t & 1));
        tc_commit(bar_empty1(s));      // ring-1 slot is free once MMA1 has read it
      };
      mbar_wait(bar_q, 0);
      tc_fence_after();
#ifdef MSCL_TC_TIMELINE
      TL(28);
      mbar_wait(bar_full1(0), 0);
      TL(29);
#endif
      if (nt > 0) issue_mma1(0);
      for (int t = 0; t < nt; ++t) {
#ifdef MSCL_TC_TIMELINE
        if (t == 3) TL(16);
#endif
        if (t + 1 < nt) issue_mma1(t + 1);
#ifdef MSCL_TC_TIMELINE
        if (t == 3) TL(17);
#endif
        if (GRAD) {
          mbar_wait(bar_pfull(t & 1), (uint32_t)(t >> 1) & 1u);
          tc_fence_after();
#ifdef MSCL_TC_TIMELINE
          if (t == 3) TL(18);
#endif
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int u = 2 * t + h;
            const int s = u % kStages2;
            mbar_wait(bar_full2(s), (uint32_t)(u / kStages2) & 1u);
            tc_fence_after();
            const uint32_t a = tmem + kColS + (uint32_t)(t & 1) * kTile + h * kHalf;
            const uint32_t lo = lo2_base + (uint32_t)s * (kW2Bytes >> 4);
            // B = the ring-2 copy read MN-major (SWIZZLE_128B_BASE32B): 8 keys per step = two 4-row atoms
            // 512 bytes apart (SBO); channel blocks kW2Slab bytes apart (LBO)
#ifndef MSCL_EXP_NOMMA2
            if (u == 0) {
#pragma unroll
              for (int j = 0; j < kHalf / 8; ++j)
                mma_ts_lh(tmem + kColO, a + j * 8, lo + ((j * 1024) >> 4), kHi2, kIdesc2, j ? 1u : 0u);
            } else {
#pragma unroll
              for (int j = 0; j < kHalf / 8; ++j)
                mma_ts_lh(tmem + kColO, a + j * 8, lo + ((j * 1024) >> 4), kHi2, kIdesc2, 1u);
            }
#endif
            tc_commit(bar_empty2(s));
          }
#ifdef MSCL_TC_TIMELINE
          if (t == 3) TL(19);
#endif
        } else {
          // no second GEMM; p_full doubles as "S consumed" (the S buffer may be overwritten by MMA1(t+2))
          mbar_wait(bar_pfull(t & 1), (uint32_t)(t >> 1) & 1u);
        }
      }
      if (GRAD) tc_commit(bar_ofull);
    }
    __syncwarp();
  } else {
    // ===================== softmax / epilogue warps (8: two per TMEM lane quarter) =====================
    const int sw = warp - 2;                      // 0..7
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = sw >> 2;                     // which 32 of a tile's 64 keys (and which half of Q / O)
    const int r = quarter * 32 + lane;            // row within the CTA's block == TMEM lane
    const int row = row0 + r;
    const bool row_ok = row < M;
    const bool warp_ok = (row0 + quarter * 32) < M;   // any valid row in this warp's lane quarter
    float shift2 = 0.f, pos2 = INFINITY;
    int64_t dup_local = -1;   // queue slot (in this shard) holding a copy of the row's positive key
    pdl_wait();               // qpack and dscale come from the prep kernel this grid may have overlapped with
    if (row_ok) {
      const float4 x = __ldg(reinterpret_cast<const float4 *>(qpack + (int64_t)row * kLd + kC));
      shift2 = x.y;
      pos2 = x.x;
      const int dup = __float_as_int(x.z);
      if (dup >= 0) dup_local = (int64_t)dup - shard_begin;
    }
    const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
    {   // this row of Q: smem (TMA, swizzled) -> TMEM columns [kColQ + 64*half, +64): the A operand of every MMA1
      mbar_wait(bar_qload, 0);
      const uint8_t *qs = gbase + kOffW2 + kW2Bytes;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int cb = half * 2 + hh;             // channel block (slab) of 32 channels
        const uint8_t *rowp = qs + cb * kQSlab + r * 128;
        uint32_t v[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 f = *reinterpret_cast<const float4 *>(rowp + ((c ^ (r & 7)) << 4));
          v[c * 4 + 0] = __float_as_uint(f.x);
          v[c * 4 + 1] = __float_as_uint(f.y);
          v[c * 4 + 2] = __float_as_uint(f.z);
          v[c * 4 + 3] = __float_as_uint(f.w);
        }
        TC_ST32(lane_base + kColQ + cb * 32, v);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_qfree);
        mbar_arrive(bar_q);
      }
      if (threadIdx.x == 64) TL(2);
    }
    float sum = 0.f;
    int cnt = 0;
    for (int t = 0; t < nt; ++t) {
      const int b = t & 1;
      const int64_t key0 = (t_begin + t) * kTile + half * 64;
      // this warp's 64 per-key scales (L1/L2-resident, warp-uniform addresses), fetched before the wait;
      // the array is padded to a multiple of 128 so the reads never leave it
      float4 d4[16];
      if (warp_ok) {
        const float4 *dsg = reinterpret_cast<const float4 *>(dscale + key0);
#pragma unroll
        for (int j = 0; j < 16; ++j) d4[j] = __ldg(dsg + j);
      }
      mbar_wait(bar_sfull(b), (uint32_t)(t >> 1) & 1u);
      tc_fence_after();
#ifdef MSCL_TC_TIMELINE
      if (threadIdx.x == 64 && t < 8) TL(8 + t);
#endif
#ifdef MSCL_EXP_NOSOFTMAX
      if (false) {
#else
      if (warp_ok) {
#endif
        // both 32-key chunks of this warp's 64 keys are fetched from TMEM up front (one wait)
        const uint32_t taddr = lane_base + kColS + (uint32_t)b * kTile + half * 64;
        uint32_t v0[32], v1[32];
        TC_LD32(taddr, v0);
        TC_LD32(taddr + 32, v1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {          // two chunks of 32 keys
          const int64_t k0 = key0 + ch * 32;
          const int64_t left = K_local - k0;
          const int nvalid = left < 32 ? (left < 0 ? 0 : (int)left) : 32;
          const int64_t dcol = dup_local - k0;
          const bool has_dup = dcol >= 0 && dcol < 32;
          uint32_t(&v)[32] = ch ? v1 : v0;
          // warp-uniform choice (tcgen05.ld/st are .sync.aligned): slow path if any row of the warp needs it
          if (nvalid == 32 && !__any_sync(0xffffffffu, has_dup))
            softmax_half<GRAD, true>(v, d4 + ch * 8, shift2, pos2, 32, -1, sum, cnt);
          else
            softmax_half<GRAD, false>(v, d4 + ch * 8, shift2, pos2, nvalid, has_dup ? (int)dcol : -1, sum, cnt);
          if (GRAD) TC_ST32(taddr + ch * 32, v);
        }
        if (GRAD) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pfull(b));
#ifdef MSCL_TC_TIMELINE
      if (threadIdx.x == 64 && t < 8) TL(20 + t);
#endif
    }
    if (threadIdx.x == 64) TL(5);
    // combine the two column halves of each row, then one plain store per row
    if (half == 1) {
      red_smem[r] = sum;
      red_smem[kRows + r] = (float)cnt;
    }
    asm volatile("bar.sync 1, %0;" ::"r"(kSoftmaxWarps * 32) : "memory");
    float *prow = part + ((int64_t)blockIdx.x * M + row) * kLd;
    if (half == 0 && row_ok)
      *reinterpret_cast<float4 *>(prow + kC) =
          make_float4(sum + red_smem[r], (float)cnt + red_smem[kRows + r], 0.f, 0.f);
    if (GRAD) {
      mbar_wait(bar_ofull, 0);
      tc_fence_after();
      if (threadIdx.x == 64) TL(6);
      // O tile: TMEM -> registers -> stage-0 buffer (all tiles are consumed by now) in the 128B-swizzled
      // slab layout -> ONE TMA store per CTA into this CTA's slab (rows >= M are clipped by the TMA unit)
      if (warp_ok) {
        uint8_t *os = gbase + kOffW;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int cb = half * 2 + hh;
          uint32_t v[32];
          TC_LD32(lane_base + kColO + cb * 32, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          uint8_t *rowp = os + cb * kQSlab + r * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<float4 *>(rowp + ((c ^ (r & 7)) << 4)) =
                make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]),
                            __uint_as_float(v[4 * c + 2]), __uint_as_float(v[4 * c + 3]));
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, %0;" ::"r"(kSoftmaxWarps * 32) : "memory");
      if (threadIdx.x == 64) {
        tma_store_4d(&tmap_part, sW, 0, row0, 0, (int)blockIdx.x);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem may go; the grid's end publishes the data
      }
    }
  }

  if (threadIdx.x == 64) TL(7);
#ifdef MSCL_TC_TIMELINE
  if (threadIdx.x == 0) {
    TL(3);
    g_timeline[(blockIdx.y * gridDim.x + blockIdx.x) * 32 + 31] = clock64();
  }
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols)
                 : "memory");
  }
}


}  // namespace tc
}  // namespace mscl

extern "C" int mscl_infonce_partial(const float *d_qpack, int32_t M, const float *d_queue,
                                    const float *d_dscale, int64_t K_local, int64_t shard_begin,
                                    float *d_part, int32_t n_part, int32_t with_grad,
                                    mscl_stream_t stream) {
  using namespace mscl::tc;
  MSCL_CHECK_ARG(d_qpack && d_queue && d_dscale && d_part, "null pointer");
  MSCL_CHECK_ARG(M > 0 && K_local > 0, "bad M=%d K_local=%lld", M, (long long)K_local);
  MSCL_CHECK_ARG(K_local < (1ll << 31), "K_local too large for a TMA coordinate");
  MSCL_CHECK_ARG((((uintptr_t)d_qpack | (uintptr_t)d_queue | (uintptr_t)d_dscale | (uintptr_t)d_part) & 15) == 0,
                 "qpack/queue/dscale/part must be 16-byte aligned");
  const int64_t n_tiles = (K_local + kTile - 1) / kTile;
  MSCL_CHECK_ARG(n_part > 0 && n_part <= n_tiles, "n_part=%d must be in [1, %lld] (one 128-key tile per CTA at least)",
                 n_part, (long long)n_tiles);
  CUtensorMap tw, tw2, tq, tp;
  int rc = make_map(&tw, d_queue, K_local, kC, kTile);
  if (rc) return rc;
  rc = make_map(&tw2, d_queue, K_local, kC, kHalf, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  rc = make_map(&tq, d_qpack, M, kLd, kRows);
  if (rc) return rc;
  rc = make_map_part(&tp, d_part, M, n_part);
  if (rc) return rc;
  const int row_blocks = (M + kRows - 1) / kRows;
  dim3 grid((unsigned)n_part, (unsigned)row_blocks);
  cudaStream_t s = mscl::as_stream(stream);
  MSCL_CUDA(mscl::ensure_dyn_smem(infonce_tc_kernel<true>, kSmemBytes));      // per device ("weak" 9)
  MSCL_CUDA(mscl::ensure_dyn_smem(infonce_tc_kernel<false>, kSmemBytes));
  if (with_grad) {
    MSCL_CUDA(mscl::launch_pdl(infonce_tc_kernel<true>, grid, dim3(kThreads), kSmemBytes, s, tw, tw2, tq, tp, d_qpack, M,
                               d_dscale, K_local, shard_begin, d_part));
  } else {
    MSCL_CUDA(mscl::launch_pdl(infonce_tc_kernel<false>, grid, dim3(kThreads), kSmemBytes, s, tw, tw2, tq, tp, d_qpack, M,
                               d_dscale, K_local, shard_begin, d_part));
  }
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

// How many partial slabs (CTAs along the key axis) mscl_infonce_partial should be given.
extern "C" int mscl_infonce_num_partials(int32_t M, int64_t K_local, int32_t num_sms) {
  using namespace mscl::tc;
  if (M <= 0 || K_local <= 0 || num_sms <= 0) return mscl::set_err(MSCL_EINVAL, "bad M / K_local / num_sms");
  const int64_t n_tiles = (K_local + kTile - 1) / kTile;
  const int row_blocks = (M + kRows - 1) / kRows;
  int64_t gx = num_sms / row_blocks;
  if (gx < 1) gx = 1;
  if (gx > n_tiles) gx = n_tiles;
  return (int)gx;
}

#ifdef MSCL_TC_TIMELINE
extern "C" int mscl_debug_timeline(unsigned long long *host_out, int n) {
  MSCL_CUDA(cudaMemcpyFromSymbol(host_out, mscl::tc::g_timeline, sizeof(unsigned long long) * n));
  return MSCL_OK;
}
#endif
