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
#include <cuda.h>

#include "common.cuh"

namespace mscl {
namespace tc {

constexpr int kC = MSCL_DIM;          // 128 channels
constexpr int kLd = MSCL_PACK_LD;     // 132
constexpr int kRows = 128;            // query rows per CTA (UMMA M)
constexpr int kTile = 64;             // keys per stage
constexpr int kStages = 3;
constexpr int kCb = 4;                // channel blocks of 32 fp32 (one 128-byte swizzle row)
constexpr int kThreads = 192;

constexpr uint32_t kWBytes = kTile * kC * 4;          // 32768: one copy of a tile
constexpr uint32_t kWSlab = kTile * 128;              // bytes per channel block of a W tile
constexpr uint32_t kStageBytes = 2 * kWBytes;         // K-major copy + MN-major copy
constexpr uint32_t kDsBytes = kTile * 4;              // 256

// shared memory map (offsets from the 1024-aligned base)
constexpr uint32_t kOffW = 0;
constexpr uint32_t kOffDs = kOffW + kStages * kStageBytes;
constexpr uint32_t kOffBar = kOffDs + kStages * kDsBytes;
constexpr uint32_t kNumBars = 2 * kStages + 1 + 2 + 2 + 1;  // full, empty, q, s_full[2], p_full[2], o_full
constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr uint32_t kSmemUsed = kOffTmemPtr + 16;
constexpr uint32_t kSmemBytes = kSmemUsed + 1024;     // slack for manual 1024-byte alignment

constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColO = 0;
constexpr uint32_t kColS = 128;
constexpr uint32_t kColQ = 256;

// instruction descriptors (cute::UMMA::InstrDescriptor bit layout):
//  [4,6) c_format=1 (f32) | [7,10) a_format=2 (tf32) | [10,13) b_format=2 (tf32)
//  [15] a_major | [16] b_major (1 = MN-major) | [17,23) N>>3 | [24,29) M>>4
constexpr uint32_t kIdescBase = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kRows >> 4) << 24);
constexpr uint32_t kIdesc1 = kIdescBase | ((uint32_t)(kTile >> 3) << 17);               // N = 64
constexpr uint32_t kIdesc2 = kIdescBase | (1u << 16) | ((uint32_t)(kC >> 3) << 17);     // N = 128, B MN-major

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- TMA ----
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void *src, uint32_t bytes,
                                             uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// ---- tcgen05 ----
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                       uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem desc]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc,
                                       uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// smem matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 |
// version=1 <<46 | layout_type <<61  (2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type = 2) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
         ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | ((uint64_t)layout_type << 61);
}

#define TC_LD32(taddr, r)                                                                       \
  asm volatile(                                                                                 \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                 \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23," \
      "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),     \
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), \
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),           \
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),           \
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])            \
      : "r"(taddr))

#define TC_ST32(taddr, r)                                                                       \
  asm volatile(                                                                                 \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "                                          \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23," \
      "%24,%25,%26,%27,%28,%29,%30,%31};" ::"r"(r[0]),                                          \
      "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),   \
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),        \
      "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),       \
      "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),       \
      "r"(r[30]), "r"(r[31]), "r"(taddr)                                                        \
      : "memory")

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b),
               "f"(c), "f"(d)
               : "memory");
}

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

// One 64-key tile of one query row: S (TMEM) -> p, row sum, hit count, P' (TMEM, over S).
template <bool GRAD, bool FULL>
__device__ __forceinline__ void softmax_tile(uint32_t taddr0, const float *ds, float shift2, float thr,
                                             int nvalid, int dupcol, float &sum, int &cnt) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t v[32];
    const uint32_t taddr = taddr0 + h * 32;
    TC_LD32(taddr, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      const float4 d4 = *reinterpret_cast<const float4 *>(ds + h * 32 + j4 * 4);
      const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = j4 * 4 + e;
        const float tval = fmaf(__uint_as_float(v[j]), dd[e], -shift2);
        float p = ex2(tval);
        bool hit = tval > thr;
        if (!FULL) {  // tail tile (keys beyond K_local) or the tile holding this row's own positive key
          const bool ok = (h * 32 + j) < nvalid;
          p = ok ? p : 0.f;
          hit = hit && ok && (h * 32 + j) != dupcol;
        }
        sum += p;
        cnt += hit ? 1 : 0;
        v[j] = __float_as_uint(GRAD ? to_tf32_rn(p * dd[e]) : 0.f);
      }
    }
    if (GRAD) TC_ST32(taddr, v);
  }
}

template <bool GRAD>
__global__ void __launch_bounds__(kThreads, 1)
infonce_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_w2,
                  const float *__restrict__ qpack, int M, const float *__restrict__ dscale,
                  int64_t K_local, int64_t shard_begin, float *__restrict__ acc) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sW = base + kOffW;
  const uint32_t sDs = base + kOffDs;
  const uint32_t bar0 = base + kOffBar;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (kStages + s); };
  const uint32_t bar_q = bar0 + 8u * (2 * kStages);
  auto bar_sfull = [&](int b) { return bar0 + 8u * (2 * kStages + 1 + b); };
  auto bar_pfull = [&](int b) { return bar0 + 8u * (2 * kStages + 3 + b); };
  const uint32_t bar_ofull = bar0 + 8u * (2 * kStages + 5);
  volatile uint32_t *tmem_ptr_smem = reinterpret_cast<volatile uint32_t *>(gbase + kOffTmemPtr);
  const float *ds_smem = reinterpret_cast<const float *>(gbase + kOffDs);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // this CTA's tile range
  const int64_t n_tiles = (K_local + kTile - 1) / kTile;
  const int64_t t_begin = n_tiles * blockIdx.x / gridDim.x;
  const int64_t t_end = n_tiles * (blockIdx.x + 1) / gridDim.x;
  const int nt = (int)(t_end - t_begin);
  const int row0 = blockIdx.y * kRows;

  if (warp == 0 && lane == 0) {
    TL(0);
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w2) : "memory");
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    mbar_init(bar_q, 128);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_sfull(b), 1);
      mbar_init(bar_pfull(b), 128);
    }
    mbar_init(bar_ofull, 1);
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
  const uint32_t tmem = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      TL(1);
      for (int t = 0; t < nt; ++t) {
        const int s = t % kStages;
        const uint32_t use = (uint32_t)(t / kStages);
        mbar_wait(bar_empty(s), (use & 1u) ^ 1u);
        mbar_arrive_expect_tx(bar_full(s), (GRAD ? 2 * kWBytes : kWBytes) + kDsBytes);
        const int64_t key0 = (t_begin + t) * kTile;
        tma_load_3d(sW + s * kStageBytes, &tmap_w, bar_full(s), 0, (int)key0, 0);
        if (GRAD) tma_load_3d(sW + s * kStageBytes + kWBytes, &tmap_w2, bar_full(s), 0, (int)key0, 0);
        bulk_load_1d(sDs + s * kDsBytes, dscale + key0, kDsBytes, bar_full(s));
      }
      TL(4);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      auto issue_mma1 = [&](int t) {
        const int s = t % kStages;
        mbar_wait(bar_full(s), (uint32_t)(t / kStages) & 1u);
        tc_fence_after();
        const uint32_t d = tmem + kColS + (uint32_t)(t & 1) * kTile;
#pragma unroll
        for (int cb = 0; cb < kCb; ++cb) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t bd = make_desc(sW + s * kStageBytes + cb * kWSlab + ks * 32, 16, 1024);
            mma_ts(d, tmem + kColQ + cb * 32 + ks * 8, bd, kIdesc1, (cb | ks) ? 1u : 0u);
          }
        }
        tc_commit(bar_sfull(t & 1));
      };
      mbar_wait(bar_q, 0);
      tc_fence_after();
      if (nt > 0) issue_mma1(0);
      for (int t = 0; t < nt; ++t) {
        if (t + 1 < nt) issue_mma1(t + 1);
        if (GRAD) {
          const int s = t % kStages;
          mbar_wait(bar_pfull(t & 1), (uint32_t)(t >> 1) & 1u);
          tc_fence_after();
          const uint32_t a = tmem + kColS + (uint32_t)(t & 1) * kTile;
#pragma unroll
          for (int j = 0; j < kTile / 8; ++j) {
            // B = the stage's second copy read MN-major (SWIZZLE_128B_BASE32B): 8 keys per step =
            // two 4-row atoms 512 bytes apart (SBO); channel blocks kWSlab bytes apart (LBO)
            const uint64_t bd = make_desc(sW + s * kStageBytes + kWBytes + j * 1024, kWSlab, 512, 1);
            mma_ts(tmem + kColO, a + j * 8, bd, kIdesc2, (t | j) ? 1u : 0u);
          }
          tc_commit(bar_empty(s));
        } else {
          // no second GEMM: the stage is free once the softmax warps have consumed dscale
          // and MMA1 has read the tile; p_full doubles as "S consumed".
          const int s = t % kStages;
          mbar_wait(bar_pfull(t & 1), (uint32_t)(t >> 1) & 1u);
          mbar_arrive(bar_empty(s));
        }
      }
      if (GRAD) tc_commit(bar_ofull);
    }
  } else {
    // ===================== softmax / epilogue warps =====================
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;            // row within the CTA's block == TMEM lane
    const int row = row0 + r;
    const bool row_ok = row < M;
    float shift2 = 0.f, thr = INFINITY;
    int64_t dup_local = -1;   // queue slot (in this shard) holding a copy of the row's positive key
    if (row_ok) {
      const float pos2 = qpack[(int64_t)row * kLd + kC];
      shift2 = qpack[(int64_t)row * kLd + kC + 1];
      thr = pos2 - shift2;
      const int dup = __float_as_int(qpack[(int64_t)row * kLd + kC + 2]);
      if (dup >= 0) dup_local = (int64_t)dup - shard_begin;
    }
    const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
    {   // this row of Q -> TMEM columns [kColQ, kColQ+128): the A operand of every MMA1
      const float4 *qsrc = reinterpret_cast<const float4 *>(qpack + (int64_t)row * kLd);
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        uint32_t v[32];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_ok) f = __ldg(qsrc + h * 8 + j4);
          v[j4 * 4 + 0] = __float_as_uint(f.x);
          v[j4 * 4 + 1] = __float_as_uint(f.y);
          v[j4 * 4 + 2] = __float_as_uint(f.z);
          v[j4 * 4 + 3] = __float_as_uint(f.w);
        }
        TC_ST32(lane_base + kColQ + h * 32, v);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      mbar_arrive(bar_q);
      if (threadIdx.x == 64) TL(2);
    }
    float sum = 0.f;
    int cnt = 0;
    for (int t = 0; t < nt; ++t) {
      const int s = t % kStages;
      const int b = t & 1;
      mbar_wait(bar_full(s), (uint32_t)(t / kStages) & 1u);   // dscale slice visible
      mbar_wait(bar_sfull(b), (uint32_t)(t >> 1) & 1u);
      tc_fence_after();
#ifdef MSCL_TC_TIMELINE
      if (threadIdx.x == 64 && t < 12) TL(8 + t);
#endif
      const int64_t key0 = (t_begin + t) * kTile;
      const int nvalid = (K_local - key0) < kTile ? (int)(K_local - key0) : kTile;
      const float *ds = ds_smem + s * kTile;
      const int64_t dcol = dup_local - key0;
      const bool has_dup = dcol >= 0 && dcol < kTile;
      // warp-uniform choice (tcgen05.ld/st are .sync.aligned): slow path if any row of the warp needs it
      if (nvalid == kTile && !__any_sync(0xffffffffu, has_dup))
        softmax_tile<GRAD, true>(lane_base + kColS + (uint32_t)b * kTile, ds, shift2, thr, nvalid, -1, sum, cnt);
      else
        softmax_tile<GRAD, false>(lane_base + kColS + (uint32_t)b * kTile, ds, shift2, thr, nvalid,
                                  has_dup ? (int)dcol : -1, sum, cnt);
      if (GRAD) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      mbar_arrive(bar_pfull(b));
#ifdef MSCL_TC_TIMELINE
      if (threadIdx.x == 64 && t < 12) TL(20 + t);
#endif
    }
    if (threadIdx.x == 64) TL(5);
    if (row_ok) {
      atomicAdd(acc + (int64_t)row * kLd + kC, sum);
      atomicAdd(acc + (int64_t)row * kLd + kC + 1, (float)cnt);
    }
    if (GRAD && nt > 0) {
      mbar_wait(bar_ofull, 0);
      tc_fence_after();
      if (threadIdx.x == 64) TL(6);
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        uint32_t v[32];
        TC_LD32(lane_base + kColO + h * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row_ok) {
          float *dst = acc + (int64_t)row * kLd + h * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            red_add_v4(dst + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                       __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
      }
    }
  }

  if (threadIdx.x == 64) TL(7);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols)
                 : "memory");
  }
}

// ---- host side: tensor maps through the driver entry point (no link-time libcuda) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void *p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
    cudaGetLastError();
    return nullptr;
  }
  fn = (EncodeTiledFn)p;
  return fn;
}

// rows x 128 fp32 matrix with row pitch ld floats, viewed as {32, rows, 4} so that one box lands
// in shared memory as 4 channel-block slabs of [box_rows][128 B], 128-byte swizzled.
static int make_map(CUtensorMap *map, const float *ptr, int64_t rows, int ld, int box_rows,
                    CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_err(MSCL_EUNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {32, (cuuint64_t)rows, 4};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, 128};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 4};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(ptr), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_err(MSCL_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld ld=%d)",
                   (int)r, (long long)rows, ld);
  return MSCL_OK;
}

}  // namespace tc
}  // namespace mscl

extern "C" int mscl_infonce_partial(const float *d_qpack, int32_t M, const float *d_queue,
                                    const float *d_dscale, int64_t K_local, int64_t shard_begin,
                                    float *d_acc, int32_t with_grad, int32_t num_sms,
                                    mscl_stream_t stream) {
  using namespace mscl::tc;
  MSCL_CHECK_ARG(d_qpack && d_queue && d_dscale && d_acc, "null pointer");
  MSCL_CHECK_ARG(M > 0 && K_local > 0, "bad M=%d K_local=%lld", M, (long long)K_local);
  MSCL_CHECK_ARG(K_local < (1ll << 31), "K_local too large for a TMA coordinate");
  MSCL_CHECK_ARG((((uintptr_t)d_qpack | (uintptr_t)d_queue | (uintptr_t)d_dscale | (uintptr_t)d_acc) & 15) == 0,
                 "qpack/queue/dscale/acc must be 16-byte aligned");
  MSCL_CHECK_ARG(num_sms > 0, "num_sms=%d", num_sms);
  CUtensorMap tw, tw2;
  int rc = make_map(&tw, d_queue, K_local, kC, kTile);
  if (rc) return rc;
  rc = make_map(&tw2, d_queue, K_local, kC, kTile, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  const int64_t n_tiles = (K_local + kTile - 1) / kTile;
  const int row_blocks = (M + kRows - 1) / kRows;
  int64_t gx = num_sms / row_blocks;
  if (gx < 1) gx = 1;
  if (gx > n_tiles) gx = n_tiles;
  dim3 grid((unsigned)gx, (unsigned)row_blocks);
  cudaStream_t s = mscl::as_stream(stream);
  if (with_grad) {
    MSCL_CUDA(cudaFuncSetAttribute(infonce_tc_kernel<true>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    infonce_tc_kernel<true><<<grid, kThreads, kSmemBytes, s>>>(tw, tw2, d_qpack, M, d_dscale, K_local, shard_begin, d_acc);
  } else {
    MSCL_CUDA(cudaFuncSetAttribute(infonce_tc_kernel<false>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    infonce_tc_kernel<false><<<grid, kThreads, kSmemBytes, s>>>(tw, tw2, d_qpack, M, d_dscale, K_local, shard_begin, d_acc);
  }
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

#ifdef MSCL_TC_TIMELINE
extern "C" int mscl_debug_timeline(unsigned long long *host_out, int n) {
  MSCL_CUDA(cudaMemcpyFromSymbol(host_out, mscl::tc::g_timeline, sizeof(unsigned long long) * n));
  return MSCL_OK;
}
#endif
