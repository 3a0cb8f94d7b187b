// K1, single-launch form: the whole InfoNCE term -- per-row positive logit and shift, per-key age decay, the
// tensor-core pass over the queue, the cross-CTA reduction and the loss / top-k / d loss/d q epilogue -- in ONE
// kernel (infonce_tc.cu + infonce.cu need three launches and a 148-slab round trip through L2 for the same result).
// replaces moco.py:481-498, heads/moco_head.py:38-77, heads/moco_head_v2.py:38-100, losses/cross_entropy_loss.py:134-138,
// core/evaluation/accuracy.py:130-149.  sm_100a only.
//
// Mathematics as in infonce_tc.cu:  S = Q W^T (scaled per key by dscale_j), P = 2^(S - shift), O = P (dscale . W).
//
// What is different from the slab form, and why (profiles/r01_tc_timeline.txt, VERDICT r01 "weak" 1):
//  * work unit = 64 keys.  A CTA owns a contiguous range of units; it processes them as 128-key PAIR tiles (a tf32
//    tcgen05 dispatch never costs less than ~61 cycles, so N = 128 keys per MMA1 dispatch is the efficient shape) plus,
//    when its unit count is odd, one 64-key HALF tile that is requested with the first loads and processed LAST, so that
//    the drain after the last byte is half as long.  1024 units over 148 CTAs is 6.92 per CTA (max 7): the 128-key split was 3.46 (max 4), i.e. the slowest CTA
//    streamed 14 % more than the average.
//  * ring 1 = two pair slots + one dedicated half-tile slot: 160 KB of a CTA's <= 224 KB are requested before anything
//    else happens (was 128 of <= 256 KB).  Ring 2 (the MN-major copy for the second GEMM, L2 hits) has two 64-key slots.
//  * no prep launch: the raw q rows and their positives arrive by TMA (first in the TMA unit's queue), the softmax warps
//    compute pos2 / shift2 from them (warp per row, transpose-reduce), round q to tf32 on its way into TMEM, and a 12th
//    warp turns birth[] into the per-key scale 0.99999^age / T * log2(e) tile by tile.
//  * no finalize launch: the row statistics (sum-exp, hit count: 8 bytes per row and CTA) are added into one small
//    accumulator with red.global.add, and the last CTA to finish (device counter) turns them into losses, top-k flags,
//    group means and the two per-row gradient coefficients -- on the four warps that are idle by then, while the
//    softmax warps of the same CTA are still storing O.  The O partials (48 KB per CTA) are NOT reduced in the forward
//    launch: a first version added them into one accumulator with a TMA reduce-add and measured ~1 TB/s of fp32 adds
//    in L2, i.e. 7 us for the 7.5 MB of K = 65536 / M = 96 (profiles/r02_k1_fused_v0_timeline.txt).  They are stored
//    as per-CTA slabs (one TMA store each, off the critical path) and summed in a fixed order by the backward kernel
//    (mscl_infonce_bwd_slabs), which needs them only when autograd asks for dq.
//
// CTA = 384 threads, one CTA per SM:
//   warp 0      ring-1 TMA producer (queue tiles from HBM, the Q tile)
//   warp 1      tcgen05.mma issuer
//   warps 2..9  softmax / epilogue (thread <-> query row == TMEM lane; two warps per lane quarter, 64 keys each)
//   warp 10     ring-2 TMA producer
//   warp 11     per-key scale (birth -> dscale) into a 2-deep shared-memory ring
// TMEM (512 columns): O [0,128) | S/P double buffer [128,384) | Q [384,512).
#include <string.h>

#include "tc_common.cuh"

namespace mscl {
namespace tcf {
using namespace mscl::tc;

constexpr int kUnit = 64;             // keys per work unit / ring-2 slot / half tile
constexpr int kTile = 128;            // keys per pair tile
constexpr int kStages2 = 2;           // ring 2 slots (64 keys each)
constexpr int kCb = 4;                // channel blocks of 32 fp32 (one 128-byte swizzle row)
constexpr int kSoftmaxWarps = 8;
constexpr int kWarps = 4 + kSoftmaxWarps;
constexpr int kThreads = kWarps * 32;                  // 384

constexpr uint32_t kPairBytes = kTile * kC * 4;        // 65536
constexpr uint32_t kPairSlab = kTile * 128;            // bytes per channel block of a pair tile
constexpr uint32_t kUnitBytes = kUnit * kC * 4;        // 32768
constexpr uint32_t kUnitSlab = kUnit * 128;
constexpr uint32_t kQBytes = kRows * kC * 4;           // 65536: the Q tile (and the positives' tile)
constexpr uint32_t kQSlab = kRows * 128;               // bytes per channel block of it

// shared memory map (the dynamic segment is 1024-byte aligned: checked at kernel entry)
constexpr uint32_t kOffPair = 0;                               // 2 x 64 KB pair slots; the [128][132] output tile at the end
constexpr uint32_t kOffSingle = kOffPair + 2 * kPairBytes;     // 32 KB: the half tile
constexpr uint32_t kOffW2 = kOffSingle + kUnitBytes;           // ring 2: 2 x 32 KB; the Q tile before the first ring-2 load
constexpr uint32_t kOffBar = kOffW2 + kStages2 * kUnitBytes;
constexpr uint32_t kNumBars = 2 + 2 + 1 + 2 * kStages2 + 1 + 2 + 2 + 1 + 1 + 1 + 2 + 2 + 1 + 1;   // (the q_load slot is now "prologue loads issued")
constexpr int kStatCopies = 16;       // the row statistics are spread over this many accumulator copies (see the epilogue)
constexpr uint32_t kOffTmemPtr = kOffBar + 8 * kNumBars;
constexpr uint32_t kOffFlag = kOffTmemPtr + 8;
constexpr uint32_t kOffDs = kOffBar + 256;                     // [2][128] floats: per-key scales of the tile in flight
constexpr uint32_t kOffRow = kOffDs + 2 * kTile * 4;           // [128] float2 (pos2, shift2); later [2][128] sum / count
constexpr uint32_t kSmemBytes = kOffRow + kRows * 8;
static_assert(kOffFlag + 8 <= kOffDs, "barrier block overflows");
static_assert(kQBytes <= kStages2 * kUnitBytes && kQBytes <= kPairBytes, "the Q tile is staged through ring 2, the positives through pair slot 1");
static_assert(kRows * kLd * 4 <= 2 * kPairBytes, "the output tile is staged through the pair slots");
static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory a CTA may use");

constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColO = 0;
constexpr uint32_t kColS = 128;
constexpr uint32_t kColQ = 384;

constexpr uint32_t kIdescBase = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kRows >> 4) << 24);
constexpr uint32_t kIdesc1 = kIdescBase | ((uint32_t)(kTile >> 3) << 17);               // N = 128 keys
constexpr uint32_t kIdesc1H = kIdescBase | ((uint32_t)(kUnit >> 3) << 17);              // N = 64 keys (half tile)
constexpr uint32_t kIdesc2 = kIdescBase | (1u << 16) | ((uint32_t)(kC >> 3) << 17);     // N = 128 channels, B MN-major

enum : int { kFlagEarlyPrefetch = 1 };

struct Params {
  const float *q;            // FUSED: raw q [M][128];  else qpack [M][132] (rows already tf32, row info at [128..131])
  const float *kpos;         // FUSED: positives [M][128]
  const int32_t *birth;      // [K_local]
  const int64_t *qstate;     // {ptr, n_enq, ..}
  const int32_t *dup_slot;   // FUSED: [M] or null
  float *ws;                 // FUSED: [kStatCopies][M][4] (sum-exp, count, 0, 0) accumulators + the CTA counter behind them
  float *part;               // [n_part][M][132] per-CTA slabs: O | sum-exp | count | 0 | 0 (null when FUSED && !GRAD)
  float *row_loss;           // FUSED outputs: [2M] loss_i, then #{negatives above the positive}
  float *rowaux;             // FUSED: [M][4] pos2, shift2, ck, co  (dq_i = gout * (ck k_i + co sum_slabs O_i))
  float *group_out;
  int64_t K_local;
  int64_t shard_begin;
  int M;
  int rows_per_group;
  int dup_age;
  int flags;
  float inv_T;
  float key_norm_bound;
  // "epoch split" (mscl_infonce_fused_multi_x): rows [0, row_split) see the queue as it was BEFORE its last enqueue --
  // every age one lower (their q rows are scaled by pre_scale = 1 / 0.99999 on their way into TMEM), the slots
  // [rep_begin, rep_begin + rep_n) holding the rep_n keys that enqueue overwrote (saved by mscl_enqueue together with
  // their births), which one CTA (blockIdx.x == ex_cta) streams as one extra pair tile carrying the same key indices.
  // row_split == 0: no such rows.
  const int32_t *xbirth;     // [rep_n] births of the overwritten keys
  int64_t rep_begin;
  int rep_n;
  int row_split;
  int ex_cta;
  float pre_scale;
};

// Several independent InfoNCE terms ("jobs": their own queries, queue and outputs) in ONE launch.  blockIdx.y runs over
// the row blocks of all jobs, blockIdx.x over the key ranges of a job, so the jobs occupy disjoint sets of SMs and the
// fixed costs of a launch -- ramp (first bytes 2-3 us after launch), drain, finalize, launch gap: ~9 of the 15 us a
// single K = 65536 term takes -- are paid once for all of them.  MSCLWithAug.objective runs its two independent
// pre-enqueue passes (W_rgb: 96 rows, W_flow: 32 rows, recognizers/mscl.py:247-269 of the reference) this way.
constexpr int kMaxJobs = 4;
struct alignas(64) JobTable {
  CUtensorMap tmap_w[kMaxJobs];     // queue [K_local][128] fp32: 128-key box, 128B swizzle (pair tiles, K-major)
  CUtensorMap tmap_wh[kMaxJobs];    // the same, 64-key box (half tile)
  CUtensorMap tmap_w2[kMaxJobs];    // 64-key box, 32-byte-atom 128B swizzle (MN-major copy)
  CUtensorMap tmap_q[kMaxJobs];     // query rows [M][128 of ld] fp32: 128-row box, 128B swizzle (rows >= M zero-filled)
  CUtensorMap tmap_k[kMaxJobs];     // FUSED: positives [M][128], same box
  CUtensorMap tmap_x;               // job x_job's overwritten keys [rep_n][128]: 128-key box (rows >= rep_n zero-filled)
  CUtensorMap tmap_x2;              // the same, 64-key box, 32-byte-atom swizzle
  Params p[kMaxJobs];
  int n_jobs;
  int x_job;                        // the one job that carries an epoch split, or -1
  int rb_begin[kMaxJobs + 1];       // first blockIdx.y of each job (row blocks of 128 query rows)
};

#ifdef MSCL_TC_TIMELINE
__device__ unsigned long long g_timeline_f[148 * 2 * 32];
__device__ __forceinline__ unsigned long long gtime_f() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TLF(slot) g_timeline_f[(blockIdx.y * gridDim.x + blockIdx.x) * 32 + (slot)] = gtime_f()
#else
#define TLF(slot)
#endif

// 32 keys of one tile for one query row, on registers (see infonce_tc.cu::softmax_half); the 32 per-key scales come
// from shared memory (warp-uniform addresses: broadcast reads).
template <bool GRAD, bool FULL>
__device__ __forceinline__ void softmax_chunk(uint32_t (&v)[32], const float4 *ds, float shift2, float pos2, uint32_t okmask,
                                              float &sum, int &cnt) {
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
      if (!FULL) {
        const bool ok = (okmask >> j) & 1u;    // not past the end, not the positive's duplicate (added exactly by the epilogue), visible to this row
        p = ok ? p : 0.f;
        hit = ok ? hit : 0u;
      }
      s4[e] += p;
      c4[e] += hit;
      v[j] = GRAD ? __float_as_uint(p * dd[e]) + 0x1000u : 0u;    // P' = p * scale, rounded to nearest tf32
    }
  }
  sum += (s4[0] + s4[1]) + (s4[2] + s4[3]);
  cnt += (int)((c4[0] + c4[1]) + (c4[2] + c4[3]));
}

// A strong (gpu-scope, L2) load: what the last CTA reads the other CTAs' statistics with.  Their adds were performed at L2
// before their ticket increments (fence in between) and this CTA's ticket increment came after all of those, so a load
// that is served by L2 sees them; no second fence (MEMBAR.ALL.GPU + L1 invalidate: ~0.9 us on the kernel's critical
// path, twice) is spent on it.
__device__ __forceinline__ float4 ld_strong_v4(const float4 *p) {
  float4 v;
  asm volatile("ld.relaxed.gpu.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_add_v2(float *addr, float a, float b) {
  asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// The last CTA, on its four idle warps (128 threads): row statistics -> per-row loss / top-k count / gradient
// coefficients, per-group means; accumulator and counter left zero.  (formulas: infonce.cu::infonce_finalize_kernel)
//   d loss_i / d q_i = gout * [ ((p0 - 1)/T + w_dup inv_Z ln2) k_i + inv_Z ln2 sum_slabs O_i ] / rows_per_group
__device__ __forceinline__ void finalize_stats(const Params &p, int tid, unsigned *counter, float *sm_rows, int sm_cap) {
  const float inv_rows = 1.0f / (float)p.rows_per_group;
  const float dsc = exp2f((float)p.dup_age * kLog2Decay);
  for (int row = tid; row < p.M; row += 128) {
    float4 *arow = reinterpret_cast<float4 *>(p.rowaux) + row;
    float4 st[kStatCopies];
#pragma unroll
    for (int c = 0; c < kStatCopies; ++c) st[c] = ld_strong_v4(reinterpret_cast<float4 *>(p.ws) + (int64_t)c * p.M + row);
    const float4 ra = ld_strong_v4(arow);
    const int dup = p.dup_slot != nullptr ? __ldg(p.dup_slot + row) : -1;
    float sum = 0.f, cnt = 0.f, w = 0.f;
#pragma unroll
    for (int c = 0; c < kStatCopies; ++c) {
      __stcg(reinterpret_cast<float4 *>(p.ws) + (int64_t)c * p.M + row, make_float4(0.f, 0.f, 0.f, 0.f));
      sum += st[c].x;
      cnt += st[c].y;
    }
    const float pos2 = ra.x, shift2 = ra.y;
    if (dup >= 0) {      // the queue entry that IS this row's positive: exact fp32 terms (see infonce.cu)
      if (pos2 * dsc > pos2) cnt += 1.f;
      const float e_dup = exp2f(fmaf(pos2, dsc, -shift2));
      sum += e_dup;
      w = e_dup * (dsc * p.inv_T * kLog2e);
    }
    const float e0 = exp2f(pos2 - shift2);
    const float Z = e0 + sum;
    const float inv_Z = 1.0f / Z;
    const float p0 = e0 * inv_Z;
    const float co = inv_Z * kLn2 * inv_rows;
    const float ck = (p0 - 1.0f) * p.inv_T * inv_rows + w * co;
    // epoch split: a pre row's negatives were scored with q / 0.99999, so d logit_j / d q carries that factor too
    __stcg(arow, make_float4(pos2, shift2, ck, row < p.row_split ? co * p.pre_scale : co));
    const float loss = (shift2 + log2f(Z) - pos2) * kLn2;
    p.row_loss[row] = loss;
    p.row_loss[p.M + row] = cnt;
    if (2 * p.M <= sm_cap) {        // the group stage reads them back from shared memory (no second L2 round trip)
      sm_rows[row] = loss;
      sm_rows[p.M + row] = cnt;
    }
  }
#ifdef MSCL_TC_TIMELINE
  if (tid == 0) TLF(24);
#endif
  __threadfence_block();
  asm volatile("bar.sync 2, 128;" ::: "memory");
  const bool from_smem = 2 * p.M <= sm_cap;
  const int wid = tid >> 5, lane = tid & 31;
  const int n_groups = p.M / p.rows_per_group;
  for (int g = wid; g < n_groups; g += 4) {
    float sl = 0.f, s1 = 0.f, s5 = 0.f;
    for (int r = lane; r < p.rows_per_group; r += 32) {
      const int row = g * p.rows_per_group + r;
      sl += from_smem ? sm_rows[row] : __ldcg(p.row_loss + row);
      const float k = from_smem ? sm_rows[p.M + row] : __ldcg(p.row_loss + p.M + row);
      s1 += (k < 1.f) ? 1.f : 0.f;
      s5 += (k < 5.f) ? 1.f : 0.f;
    }
    sl = warp_sum(sl);
    s1 = warp_sum(s1);
    s5 = warp_sum(s5);
    if (lane == 0)
      *reinterpret_cast<float4 *>(p.group_out + g * 4) = make_float4(sl * inv_rows, s1 * inv_rows, s5 * inv_rows, 0.f);
  }
  if (tid == 0) *counter = 0u;
}

template <bool GRAD, bool FUSED>
__global__ void __launch_bounds__(kThreads, 1)
infonce_fused_kernel(const __grid_constant__ JobTable jt) {
  int job = 0;
  while (job + 1 < jt.n_jobs && (int)blockIdx.y >= jt.rb_begin[job + 1]) ++job;
  const Params &p = jt.p[job];
  const CUtensorMap &tmap_w = jt.tmap_w[job], &tmap_wh = jt.tmap_wh[job], &tmap_w2 = jt.tmap_w2[job];
  const CUtensorMap &tmap_q = jt.tmap_q[job], &tmap_k = jt.tmap_k[job];
  const unsigned n_ctas_job = gridDim.x * (unsigned)(jt.rb_begin[job + 1] - jt.rb_begin[job]);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *gbase = smem_raw;
  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) asm volatile("trap;");      // the swizzled tile layouts need a 1024-byte aligned base
  const uint32_t sPair = base + kOffPair;
  const uint32_t sSingle = base + kOffSingle;
  const uint32_t sW2 = base + kOffW2;
  const uint32_t bar0 = base + kOffBar;
  auto bar_full1 = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty1 = [&](int s) { return bar0 + 8u * (2 + s); };
  const uint32_t bar_fullH = bar0 + 8u * 4;
  auto bar_full2 = [&](int s) { return bar0 + 8u * (5 + s); };
  auto bar_empty2 = [&](int s) { return bar0 + 8u * (5 + kStages2 + s); };
  constexpr int kB = 5 + 2 * kStages2;
  const uint32_t bar_q = bar0 + 8u * kB;
  auto bar_sfull = [&](int b) { return bar0 + 8u * (kB + 1 + b); };
  auto bar_pfull = [&](int b) { return bar0 + 8u * (kB + 3 + b); };
  const uint32_t bar_ofull = bar0 + 8u * (kB + 5);
  const uint32_t bar_qload = bar0 + 8u * (kB + 6);
  const uint32_t bar_qfree = bar0 + 8u * (kB + 7);
  auto bar_dfull = [&](int b) { return bar0 + 8u * (kB + 8 + b); };
  auto bar_dfree = [&](int b) { return bar0 + 8u * (kB + 10 + b); };
  const uint32_t bar_ticket = bar0 + 8u * (kB + 12);           // this CTA's ticket (last or not) is in last_flag
  const uint32_t bar_kload = bar0 + 8u * (kB + 13);            // the positives' tile landed (TMA)
  static_assert(kB + 14 == kNumBars, "barrier count");
  volatile uint32_t *tmem_ptr_smem = reinterpret_cast<volatile uint32_t *>(gbase + kOffTmemPtr);
  volatile int *last_flag = reinterpret_cast<volatile int *>(gbase + kOffFlag);
  float *ds_smem = reinterpret_cast<float *>(gbase + kOffDs);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;

  // this CTA's schedule: units [u_begin, u_end) of 64 keys -> an optional half tile (first) + pair tiles
  const int64_t n_units = (p.K_local + kUnit - 1) / kUnit;
  const int64_t u_begin = n_units * blockIdx.x / gridDim.x;
  const int64_t u_end = n_units * (blockIdx.x + 1) / gridDim.x;
  const int nu = (int)(u_end - u_begin);
  const int hasH = nu & 1;
  const int npn = nu >> 1;                          // pair tiles of the queue
  // epoch split: the one CTA (per row block) that also streams the keys the last enqueue overwrote, as one more pair tile
  const bool job_x = FUSED && job == jt.x_job;
  const int hasX = (job_x && (int)blockIdx.x == p.ex_cta) ? 1 : 0;
  const int n_xu = (p.rep_n + kUnit - 1) / kUnit;   // 64-key units of it that hold keys (1 or 2)
  const int np = npn + hasX;                        // pair steps (ring-1 pair slots)
  const int nt = np + hasH;                         // processing steps
  const int64_t key_begin = u_begin * kUnit;
  const int64_t key_end = u_end * kUnit < p.K_local ? u_end * kUnit : p.K_local;
  const int64_t rep_end = p.rep_begin + p.rep_n;
  const int row0 = ((int)blockIdx.y - jt.rb_begin[job]) * kRows;
  // step i covers keys [key0(i), key0(i) + (half tile ? 64 : 128))
  // processing order = key order: the pair tiles first (the overwritten keys' tile after them), the half tile (when the unit count
  // is odd) LAST -- it is requested with the first loads and waits in its own slot; as the last step it halves the drain
  // (MMA1 -> softmax -> MMA2 of 64 keys instead of 128 after the last byte has landed)
  auto step_is_h = [&](int i) { return hasH && i == nt - 1; };
  auto step_is_x = [&](int i) { return hasX && i == npn; };
  auto step_key0 = [&](int i) {
    return step_is_x(i) ? p.rep_begin : key_begin + (int64_t)(step_is_h(i) ? npn : i) * kTile;
  };
  auto step_units = [&](int i) { return step_is_h(i) ? 1 : (step_is_x(i) ? n_xu : 2); };

  if (warp == 0 && lane == 0) {
    TLF(0);
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_q) : "memory");
    if (FUSED) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_wh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w2) : "memory");
    if (hasX) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&jt.tmap_x) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&jt.tmap_x2) : "memory");
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_full1(s), 1);
      mbar_init(bar_empty1(s), 1);
    }
    mbar_init(bar_fullH, 1);
    for (int s = 0; s < kStages2; ++s) {
      mbar_init(bar_full2(s), 1);
      mbar_init(bar_empty2(s), 1);
    }
    mbar_init(bar_q, kSoftmaxWarps);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_sfull(b), 1);
      mbar_init(bar_pfull(b), kSoftmaxWarps);
      mbar_init(bar_dfull(b), 1);
      mbar_init(bar_dfree(b), kSoftmaxWarps);
    }
    mbar_init(bar_ofull, 1);
    mbar_init(bar_qload, 1);
    mbar_init(bar_qfree, kSoftmaxWarps);
    mbar_init(bar_ticket, 1);
    mbar_init(bar_kload, 1);
    *last_flag = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + kOffTmemPtr),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);

  if (warp == 0) {
    // ===================== ring-1 producer =====================
    if (elect_one()) {
      TLF(1);
      auto load_pair = [&](int pi) {
        const int s = pi & 1;
        mbar_wait(bar_empty1(s), ((uint32_t)(pi >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(bar_full1(s), kPairBytes);
        if (hasX && pi == npn)      // the overwritten keys (rows >= rep_n of the box are zero-filled and count as landed bytes)
          tma_load_3d(sPair + s * kPairBytes, &jt.tmap_x, bar_full1(s), 0, 0, 0);
        else
          tma_load_3d(sPair + s * kPairBytes, &tmap_w, bar_full1(s), 0, (int)(key_begin + (int64_t)pi * kTile), 0);
      };
      auto load_half = [&]() {
        if (hasH) {
          mbar_arrive_expect_tx(bar_fullH, kUnitBytes);
          tma_load_3d(sSingle, &tmap_wh, bar_fullH, 0, (int)(key_begin + (int64_t)npn * kTile), 0);
        }
      };
      // kFlagEarlyPrefetch: the caller guarantees the queue was not written by the launch this grid may overlap with
      // (programmatic dependent launch), so its tiles may be requested before the dependency wait.
      // The query rows (and, FUSED, their positives) every CTA needs come FIRST in the TMA unit's queue after the wait:
      // per-thread loads of those 96 KB were throttled by the load/store unit's miss capacity (1.3 us just to issue
      // them) and, behind 160 KB of tile requests, took ~4 us to arrive.  Q is staged in the ring-2 area, the positives
      // in pair slot 1 -- whose first queue tile therefore waits until the softmax warps have consumed them (bar_qfree).
      // Request order of the queue tiles = processing order: pair 0, the half tile's slot (processed last, but its
      // request does not have to wait for anything), pair 1, ...
      // (never the overwritten keys' tile: the enqueue that saved them may be the launch before this one)
      const bool early = (p.flags & kFlagEarlyPrefetch) != 0 && npn > 0;
      if (early) load_pair(0);
      pdl_wait();
      pdl_trigger();         // only after the wait: a dependent of THIS grid may then assume this grid's predecessors are done
      mbar_arrive_expect_tx(bar_qload, kQBytes);
      tma_load_3d(sW2, &tmap_q, bar_qload, 0, row0, 0);           // MMA1 needs Q: first
      if (!early && np > 0) load_pair(0);                         // ... and the first queue tile: second
      if (FUSED) {                                                 // the positives are only needed by the first softmax
        mbar_arrive_expect_tx(bar_kload, kQBytes);
        tma_load_3d(sPair + kPairBytes, &tmap_k, bar_kload, 0, row0, 0);
      }
      load_half();
      if (np > 1) {
        if (FUSED) mbar_wait(bar_qfree, 0);
        load_pair(1);
      }
      TLF(18);
      for (int pi = 2; pi < np; ++pi) load_pair(pi);
      TLF(4);
    }
    __syncwarp();
  } else if (warp == 2 + kSoftmaxWarps) {
    // ===================== ring-2 producer: the MN-major copy of every unit (the same bytes again: L2 hits) ======
    if (GRAD && elect_one()) {
      if (!(p.flags & kFlagEarlyPrefetch)) pdl_wait();
      mbar_wait(bar_qfree, 0);                      // ring 2 held the Q tile
      int u = 0;
      for (int i = 0; i < nt; ++i) {
        const bool is_x = step_is_x(i);
        const int n_un = step_units(i);
        const int64_t k0 = step_key0(i);
        for (int h = 0; h < n_un; ++h, ++u) {
          const int s = u % kStages2;
          mbar_wait(bar_empty2(s), ((uint32_t)(u / kStages2) & 1u) ^ 1u);
          mbar_arrive_expect_tx(bar_full2(s), kUnitBytes);
          if (is_x)
            tma_load_3d(sW2 + s * kUnitBytes, &jt.tmap_x2, bar_full2(s), 0, h * kUnit, 0);
          else
            tma_load_3d(sW2 + s * kUnitBytes, &tmap_w2, bar_full2(s), 0, (int)(k0 + (int64_t)h * kUnit), 0);
        }
      }
    }
#ifdef MSCL_TC_TIMELINE
    else if (lane == 31 && nt > 0) {      // an idle lane watches the first MMA1 complete
      mbar_wait(bar_sfull(0), 0);
      TLF(15);
    }
#endif
    __syncwarp();
  } else if (warp == 3 + kSoftmaxWarps) {
    // ===================== per-key scale: 0.99999^(n_enq - birth_j) / T * log2(e), 0 for keys outside the tile ======
    pdl_wait();
    const int64_t n_enq = p.qstate[1];
    const float sc = p.inv_T * kLog2e;
    for (int i = 0; i < nt; ++i) {
      const int b = i & 1;
      mbar_wait(bar_dfree(b), ((uint32_t)(i >> 1) & 1u) ^ 1u);
      const bool is_x = step_is_x(i);
      const int64_t k0 = step_key0(i) + 4 * lane;
      const int64_t tile_end = step_key0(i) + (step_is_h(i) ? kUnit : kTile);
      const int64_t lim_q = tile_end < key_end ? tile_end : key_end;
      const int64_t lim = is_x ? rep_end : lim_q;
      int bi[4] = {0, 0, 0, 0};
      if (is_x) {       // the overwritten keys: the age they would have now; the rows that see them carry the 1 / 0.99999
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (k0 + e < lim) bi[e] = __ldg(p.xbirth + (k0 + e - p.rep_begin));
      } else if (k0 + 3 < lim) {
        const int4 t = __ldg(reinterpret_cast<const int4 *>(p.birth + k0));
        bi[0] = t.x; bi[1] = t.y; bi[2] = t.z; bi[3] = t.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (k0 + e < lim) bi[e] = __ldg(p.birth + k0 + e);
      }
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        v[e] = (k0 + e < lim) ? exp2f((float)(n_enq - (int64_t)bi[e]) * kLog2Decay) * sc : 0.f;
      reinterpret_cast<float4 *>(ds_smem + b * kTile)[lane] = make_float4(v[0], v[1], v[2], v[3]);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_dfull(b));
#ifdef MSCL_TC_TIMELINE
      if (lane == 0 && i == 0) TLF(19);
#endif
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t kHi1 = (1024u >> 4) | (1u << 14) | (2u << 29);      // K-major, SWIZZLE_128B, SBO 1024
    constexpr uint32_t kHi2 = (512u >> 4) | (1u << 14) | (1u << 29);       // MN-major, SWIZZLE_128B_BASE32B, SBO 512
    const uint32_t lo1_pair = ((sPair & 0x3FFFFu) >> 4) | ((16u >> 4) << 16);
    const uint32_t lo1_single = ((sSingle & 0x3FFFFu) >> 4) | ((16u >> 4) << 16);
    const uint32_t lo2_base = ((sW2 & 0x3FFFFu) >> 4) | ((kUnitSlab >> 4) << 16);
    if (elect_one()) {
      auto issue_mma1 = [&](int i) {
        const uint32_t d = tmem + kColS + (uint32_t)(i & 1) * kTile;
        if (step_is_h(i)) {
          mbar_wait(bar_fullH, 0);
          tc_fence_after();
#pragma unroll
          for (int cb = 0; cb < kCb; ++cb) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              mma_ts_lh(d, tmem + kColQ + cb * 32 + ks * 8, lo1_single + ((cb * kUnitSlab + ks * 32) >> 4), kHi1, kIdesc1H,
                        (cb | ks) ? 1u : 0u);
          }
          tc_commit(bar_sfull(i & 1));
        } else {
          const int pi = i;
          const int s = pi & 1;
          mbar_wait(bar_full1(s), (uint32_t)(pi >> 1) & 1u);
          tc_fence_after();
          const uint32_t lo = lo1_pair + (uint32_t)s * (kPairBytes >> 4);
#pragma unroll
          for (int cb = 0; cb < kCb; ++cb) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              mma_ts_lh(d, tmem + kColQ + cb * 32 + ks * 8, lo + ((cb * kPairSlab + ks * 32) >> 4), kHi1, kIdesc1,
                        (cb | ks) ? 1u : 0u);
          }
          tc_commit(bar_sfull(i & 1));
          tc_commit(bar_empty1(s));      // the pair slot is free once MMA1 has read it
        }
      };
      mbar_wait(bar_q, 0);
      tc_fence_after();
      TLF(28);
      if (nt > 0) issue_mma1(0);
      TLF(29);
      int u_run = 0;                     // 64-key units handed to MMA2 so far (ring-2 slot and phase)
      for (int i = 0; i < nt; ++i) {
        if (i + 1 < nt) issue_mma1(i + 1);
#ifdef MSCL_TC_TIMELINE
        if (i < 2) TLF(30 + i);
#endif
        mbar_wait(bar_pfull(i & 1), (uint32_t)(i >> 1) & 1u);
        tc_fence_after();
        if (GRAD) {
          const int n_un = step_units(i);
          for (int h = 0; h < n_un; ++h, ++u_run) {
            const int u = u_run;
            const int s = u % kStages2;
            mbar_wait(bar_full2(s), (uint32_t)(u / kStages2) & 1u);
            tc_fence_after();
            const uint32_t a = tmem + kColS + (uint32_t)(i & 1) * kTile + h * kUnit;
            const uint32_t lo = lo2_base + (uint32_t)s * (kUnitBytes >> 4);
            // B = the ring-2 copy read MN-major (SWIZZLE_128B_BASE32B): 8 keys per step = two 4-row atoms
            // 512 bytes apart (SBO); channel blocks kUnitSlab bytes apart (LBO)
            if (u == 0) {
#pragma unroll
              for (int j = 0; j < kUnit / 8; ++j)
                mma_ts_lh(tmem + kColO, a + j * 8, lo + ((j * 1024) >> 4), kHi2, kIdesc2, j ? 1u : 0u);
            } else {
#pragma unroll
              for (int j = 0; j < kUnit / 8; ++j)
                mma_ts_lh(tmem + kColO, a + j * 8, lo + ((j * 1024) >> 4), kHi2, kIdesc2, 1u);
            }
            tc_commit(bar_empty2(s));
          }
        }
      }
      if (GRAD) tc_commit(bar_ofull);
    }
    __syncwarp();
  } else {
    // ===================== softmax / epilogue warps (8: two per TMEM lane quarter) =====================
    const int sw = warp - 2;                      // 0..7
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = sw >> 2;                     // which 64 of a pair tile's 128 keys (and which half of Q / O)
    const int r = quarter * 32 + lane;            // row within the CTA's block == TMEM lane
    const int row = row0 + r;
    const bool row_ok = row < p.M;
    const bool warp_ok = (row0 + quarter * 32) < p.M;
    float shift2 = 0.f, pos2 = INFINITY;
    int64_t dup_local = -1;
    pdl_wait();
    const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
    float2 *rowinfo = reinterpret_cast<float2 *>(gbase + kOffRow);
    const float sc = p.inv_T * kLog2e;
    // ---- Q tile (TMA, 4 channel-block slabs of [128 rows][128 B], 128B-swizzled; rows >= M zero) -> tf32 -> TMEM, thread
    // per row; FUSED: pos2 = q.k / T * log2e and shift2 = |q| * bound / T * log2e from the Q and positives tiles, warp per row
    const uint8_t *qs = gbase + kOffW2;
    const uint8_t *ks = gbase + kOffPair + kPairBytes;
    constexpr int kRW = kRows / kSoftmaxWarps;     // 16 rows per warp
    mbar_wait(bar_qload, 0);
    const bool is_pre = FUSED && row < p.row_split;       // sees the queue as it was one enqueue ago (epoch split)
    const float qfac = is_pre ? p.pre_scale : 1.f;
    {   // this row of Q -> TMEM columns [kColQ + 64*half, +64): the A operand of every MMA1
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int cb = half * 2 + hh;
        const uint8_t *rowp = qs + cb * kQSlab + r * 128;
        uint32_t v[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 f = *reinterpret_cast<const float4 *>(rowp + ((c ^ (r & 7)) << 4));
          if (FUSED) f = to_tf32_rn(make_float4(f.x * qfac, f.y * qfac, f.z * qfac, f.w * qfac));
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
      if (lane == 0) mbar_arrive(bar_q);
      if (threadIdx.x == 64) TLF(2);
    }
    if (FUSED) {
      mbar_wait(bar_kload, 0);
      // vals[u] = this lane's part of q.k of row slot u (row sw + 8u), vals[16 + u] = of |q|^2; lane <-> 16-byte chunk
      float vals[2 * kRW];
      const int slab = lane >> 3, cpos = lane & 7;
#pragma unroll
      for (int u = 0; u < kRW; ++u) {
        const int rl = sw + u * kSoftmaxWarps;
        const int off = slab * kQSlab + rl * 128 + ((cpos ^ (rl & 7)) << 4);
        const float4 a = *reinterpret_cast<const float4 *>(qs + off);
        const float4 b = *reinterpret_cast<const float4 *>(ks + off);
        vals[u] = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
        vals[kRW + u] = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
      }
      if (threadIdx.x == 64) TLF(12);
      // transpose-reduce: 32 values x 32 lanes -> lane l ends up with the warp total of value l, in 31 shuffles (a
      // butterfly per value would be 160); at each level a lane hands over the half of its values its partner keeps
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) {
        const bool up = lane & s;
#pragma unroll
        for (int i = 0; i < s; ++i) {
          const float lo = vals[i], hi = vals[i + s];
          const float recv = __shfl_xor_sync(0xffffffffu, up ? lo : hi, s);
          vals[i] = (up ? hi : lo) + recv;
        }
      }
      // lane u < 16 holds q.k of row slot u, lane 16 + u its |q|^2
      const float ss = __shfl_down_sync(0xffffffffu, vals[0], 16);
      const int rl = sw + lane * kSoftmaxWarps;
      if (lane < kRW && row0 + rl < p.M) {
        const float2 ri = make_float2(vals[0] * sc, sqrtf(ss) * p.key_norm_bound * sc);
        rowinfo[rl] = ri;
        // one CTA per row block publishes the pair for the finalising CTA
        if (blockIdx.x == 0) *reinterpret_cast<float2 *>(p.rowaux + (int64_t)(row0 + rl) * 4) = ri;
      }
    }
    if (threadIdx.x == 64) TLF(13);
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_qfree);         // the staging areas go back to ring 2 / ring 1
    if (FUSED) {
      asm volatile("bar.sync 1, %0;" ::"r"(kSoftmaxWarps * 32) : "memory");
      if (row_ok) {
        const float2 ri = rowinfo[r];
        pos2 = ri.x;
        shift2 = ri.y;
        if (p.dup_slot != nullptr) {
          const int dup = __ldg(p.dup_slot + row);
          if (dup >= 0) dup_local = (int64_t)dup - p.shard_begin;
        }
      }
    } else if (row_ok) {
      const float4 x = __ldg(reinterpret_cast<const float4 *>(p.q + (int64_t)row * kLd + kC));
      shift2 = x.y;
      pos2 = x.x;
      const int dup = __float_as_int(x.z);
      if (dup >= 0) dup_local = (int64_t)dup - p.shard_begin;
    }
    float sum = 0.f;
    int cnt = 0;
    if (threadIdx.x == 64) TLF(14);
    for (int i = 0; i < nt; ++i) {
      const int b = i & 1;
      const bool is_h = step_is_h(i);
      // a pair tile: this warp's 64 keys as two 32-key chunks; the half tile: one 32-key chunk per warp
      const bool is_x = step_is_x(i);
      const int koff = is_h ? half * 32 : half * kUnit;
      const int nch = is_h ? 1 : 2;
      const int64_t key0 = step_key0(i) + koff;
      const int64_t lim = is_x ? rep_end : key_end;
      const bool active = warp_ok;
      mbar_wait(bar_dfull(b), (uint32_t)(i >> 1) & 1u);
      mbar_wait(bar_sfull(b), (uint32_t)(i >> 1) & 1u);
      tc_fence_after();
#ifdef MSCL_TC_TIMELINE
      if (threadIdx.x == 64 && i < 8) TLF(8 + i);
#endif
      if (active) {
        const uint32_t taddr = lane_base + kColS + (uint32_t)b * kTile + koff;
        const float4 *ds = reinterpret_cast<const float4 *>(ds_smem + b * kTile + koff);
        uint32_t v0[32], v1[32];
        TC_LD32(taddr, v0);
        if (nch == 2) TC_LD32(taddr + 32, v1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          if (ch >= nch) break;                    // warp-uniform
          const int64_t k0 = key0 + ch * 32;
          const int64_t left = lim - k0;
          const int nvalid = left < 32 ? (left < 0 ? 0 : (int)left) : 32;
          const int64_t dcol = dup_local - k0;
          const bool has_dup = dcol >= 0 && dcol < 32;
          // epoch split: the slots the last enqueue wrote are invisible to the pre rows in the queue's own tiles; the
          // overwritten keys' tile is visible to the pre rows only
          const bool split_here = job_x && (is_x || (k0 < rep_end && k0 + 32 > p.rep_begin));
          uint32_t(&v)[32] = ch ? v1 : v0;
          // warp-uniform choice (tcgen05.ld/st are .sync.aligned): slow path if any row of the warp needs it
          if (nvalid == 32 && !split_here && !__any_sync(0xffffffffu, has_dup)) {
            softmax_chunk<GRAD, true>(v, ds + ch * 8, shift2, pos2, 0xffffffffu, sum, cnt);
          } else {
            uint32_t ok = nvalid >= 32 ? 0xffffffffu : ((1u << nvalid) - 1u);
            if (has_dup) ok &= ~(1u << (int)dcol);
            if (split_here) {
              if (is_x) {
                ok = is_pre ? ok : 0u;
              } else if (is_pre) {
                const int64_t lo = p.rep_begin - k0, hi = rep_end - k0;
                const int l = lo < 0 ? 0 : (int)lo, h = hi > 32 ? 32 : (int)hi;      // 0 <= l < h <= 32 here
                const uint32_t below_h = h >= 32 ? 0xffffffffu : ((1u << h) - 1u);
                ok &= ~(below_h & ~((1u << l) - 1u));
              }
            }
            softmax_chunk<GRAD, false>(v, ds + ch * 8, shift2, pos2, ok, sum, cnt);
          }
          if (GRAD) TC_ST32(taddr + ch * 32, v);
        }
        if (GRAD) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_pfull(b));
        mbar_arrive(bar_dfree(b));
      }
#ifdef MSCL_TC_TIMELINE
      if (threadIdx.x == 64 && i < 8) TLF(20 + i);
#endif
    }
    if (threadIdx.x == 64) TLF(5);
    // combine the two key halves of each row (the row-info array is dead: every thread holds its pos2 / shift2)
    float *red = reinterpret_cast<float *>(gbase + kOffRow);
    asm volatile("bar.sync 1, %0;" ::"r"(kSoftmaxWarps * 32) : "memory");
    if (half == 1) {
      red[r] = sum;
      red[kRows + r] = (float)cnt;
    }
    asm volatile("bar.sync 1, %0;" ::"r"(kSoftmaxWarps * 32) : "memory");
    if (half == 0) {
      sum += red[r];
      cnt += (int)red[kRows + r];
    }
    if (FUSED) {
      // hand the row statistics to the scale warp (idle by now): it adds them into the accumulator, takes this CTA's
      // ticket and finalises if it is the last -- none of which the O epilogue below has to wait for
      if (half == 0) {
        red[r] = sum;
        red[kRows + r] = (float)cnt;
      }
      asm volatile("bar.arrive 3, %0;" ::"r"(kSoftmaxWarps * 32 + 32) : "memory");
    }
    if (GRAD) {
      // [128][132] output tile (O | sum | count | 0 | 0) in the pair slots (every tile is consumed once o_full fires):
      // thread <-> row out of TMEM, then warp <-> row out to this CTA's slab, 512 contiguous bytes per store instruction
      // (16-byte stores straight from the row-per-thread registers are 3072 separate L2 requests per CTA: 2 us;
      //  one TMA store of the tile: 1.5 us at the ~85 GB/s one SM's TMA unit moves)
      float *tile = reinterpret_cast<float *>(gbase + kOffPair);
      if (nt > 0) {
        mbar_wait(bar_ofull, 0);
        tc_fence_after();
      }
      if (threadIdx.x == 64) TLF(6);
      if (warp_ok) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int cb = half * 2 + hh;
          uint32_t v[32];
          if (nt > 0) {
            TC_LD32(lane_base + kColO + cb * 32, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) v[c] = 0u;
          }
          float4 *dst = reinterpret_cast<float4 *>(tile + r * kLd + cb * 32);
#pragma unroll
          for (int c = 0; c < 8; ++c)
            dst[c] = make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]), __uint_as_float(v[4 * c + 2]),
                                 __uint_as_float(v[4 * c + 3]));
        }
        if (half == 0) *reinterpret_cast<float4 *>(tile + r * kLd + kC) = make_float4(sum, (float)cnt, 0.f, 0.f);
      }
      asm volatile("bar.sync 1, %0;" ::"r"(kSoftmaxWarps * 32) : "memory");
      const int n_rows = p.M - row0 < kRows ? p.M - row0 : kRows;
      float *slab = p.part + ((int64_t)blockIdx.x * p.M + row0) * kLd;
      for (int rl = sw; rl < n_rows; rl += kSoftmaxWarps) {
        const float4 *src = reinterpret_cast<const float4 *>(tile + rl * kLd);
        float4 *dst = reinterpret_cast<float4 *>(slab + (int64_t)rl * kLd);
        stg_stream(dst + lane, src[lane]);
        if (lane == 0) stg_stream(dst + 32, src[32]);
      }
    } else if (!FUSED) {
      float *prow = p.part + ((int64_t)blockIdx.x * p.M + row) * kLd;
      if (half == 0 && row_ok) *reinterpret_cast<float4 *>(prow + kC) = make_float4(sum, (float)cnt, 0.f, 0.f);
    }
    if (threadIdx.x == 64) TLF(16);
  }

  if (FUSED && warp == 3 + kSoftmaxWarps) {
    // row statistics -> one of kStatCopies small accumulators (148 same-address atomics serialise in L2 at ~13 ns each:
    // 2 us when every CTA hits the same 2 floats per row; 16 copies make it ~10 adds per address), then this CTA's ticket
    const float *red = reinterpret_cast<const float *>(gbase + kOffRow);
    asm volatile("bar.sync 3, %0;" ::"r"(kSoftmaxWarps * 32 + 32) : "memory");
    if (nt > 0) {
#pragma unroll
      for (int u = 0; u < kRows / 32; ++u) {
        const int rl = lane + 32 * u;
        if (row0 + rl < p.M) {
          float *wrow = p.ws + ((int64_t)(blockIdx.x % kStatCopies) * p.M + row0 + rl) * 4;
          red_add_v2(wrow, red[rl], red[kRows + rl]);
        }
      }
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");      // the adds above are performed before the ticket below
    __syncwarp();
    if (lane == 0) {
      TLF(25);
      unsigned *counter = reinterpret_cast<unsigned *>(p.ws + (int64_t)kStatCopies * p.M * 4);
      const unsigned done = atomicAdd(counter, 1u);
#ifdef MSCL_TC_TIMELINE
      if (done != 0xffffffffu) TLF(26);
#endif
      // (no acquire fence here: the last CTA reads the statistics with strong L2 loads, see ld_strong_v4)
      *last_flag = (done == n_ctas_job - 1u) ? 2 : 1;
      mbar_arrive(bar_ticket);                   // release: the flag is visible to whoever completes the wait below
      TLF(7);
    }
  }
  if (FUSED && (warp < 2 || warp >= 2 + kSoftmaxWarps)) {
    // the four warps that are idle once their role is done: wait for this CTA's ticket; the last CTA finalises here,
    // concurrently with its softmax warps storing O
    mbar_wait(bar_ticket, 0);
    const int f = *last_flag;
    if (f == 2) {
      if (warp == 0 && lane == 0) TLF(27);
      const int tid = (warp < 2 ? warp : warp - kSoftmaxWarps) * 32 + lane;
      // scratch for the row losses: the half-tile slot (consumed long ago; the softmax warps stage O in the pair slots)
      finalize_stats(p, tid, reinterpret_cast<unsigned *>(p.ws + (int64_t)kStatCopies * p.M * 4),
                     reinterpret_cast<float *>(gbase + kOffSingle), (int)(kUnitBytes / 4));
      if (tid == 0) TLF(17);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
  }
  if (threadIdx.x == 0) TLF(3);
}

static int launch(bool fused, bool grad, int n_jobs, const Params *ps, const float *const *queues, int n_part, cudaStream_t s,
                  int x_job = -1, const float *d_xkeys = nullptr) {
  MSCL_CHECK_ARG(n_jobs >= 1 && n_jobs <= kMaxJobs, "n_jobs=%d must be in [1, %d]", n_jobs, kMaxJobs);
  MSCL_CHECK_ARG(n_part > 0, "n_part=%d must be positive", n_part);
  JobTable jt;
  memset(&jt, 0, sizeof(jt));
  jt.n_jobs = n_jobs;
  jt.x_job = x_job;
  int64_t max_units = 0;
  for (int j = 0; j < n_jobs; ++j) {
    const Params &p = ps[j];
    const int64_t n_units = (p.K_local + kUnit - 1) / kUnit;
    max_units = n_units > max_units ? n_units : max_units;
    int rc = make_map(&jt.tmap_w[j], queues[j], p.K_local, kC, kTile);
    if (rc) return rc;
    rc = make_map(&jt.tmap_wh[j], queues[j], p.K_local, kC, kUnit);
    if (rc) return rc;
    rc = make_map(&jt.tmap_w2[j], queues[j], p.K_local, kC, kUnit, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
    rc = make_map(&jt.tmap_q[j], p.q, p.M, fused ? kC : kLd, kRows);
    if (rc) return rc;
    if (fused) {
      rc = make_map(&jt.tmap_k[j], p.kpos, p.M, kC, kRows);
      if (rc) return rc;
    }
    jt.p[j] = p;
    jt.rb_begin[j + 1] = jt.rb_begin[j] + (p.M + kRows - 1) / kRows;
    if (j == x_job) {
      rc = make_map(&jt.tmap_x, d_xkeys, p.rep_n, kC, kTile);
      if (rc) return rc;
      rc = make_map(&jt.tmap_x2, d_xkeys, p.rep_n, kC, kUnit, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
      if (rc) return rc;
      // the extra pair tile goes to the CTA that streams the fewest units of the queue (the first of them)
      int best = 0;
      int64_t best_nu = -1;
      for (int b = 0; b < n_part; ++b) {
        const int64_t nu = n_units * (b + 1) / n_part - n_units * b / n_part;
        if (best_nu < 0 || nu < best_nu) {
          best_nu = nu;
          best = b;
        }
      }
      jt.p[j].ex_cta = best;
    }
  }
  for (int j = n_jobs; j < kMaxJobs; ++j) jt.rb_begin[j + 1] = jt.rb_begin[n_jobs];
  MSCL_CHECK_ARG(n_part <= max_units, "n_part=%d exceeds the %lld 64-key units of the largest queue", n_part, (long long)max_units);
  MSCL_CHECK_ARG(jt.rb_begin[n_jobs] <= 65535, "too many row blocks");
  dim3 grid((unsigned)n_part, (unsigned)jt.rb_begin[n_jobs]);
#define MSCL_FUSED_LAUNCH(G, F)                                                                                \
  do {                                                                                                         \
    MSCL_CUDA(mscl::ensure_dyn_smem(infonce_fused_kernel<G, F>, kSmemBytes));                                  \
    MSCL_CUDA(mscl::launch_pdl(infonce_fused_kernel<G, F>, grid, dim3(kThreads), kSmemBytes, s, jt));          \
  } while (0)
  if (fused && grad) MSCL_FUSED_LAUNCH(true, true);
  else if (fused) MSCL_FUSED_LAUNCH(false, true);
  else if (grad) MSCL_FUSED_LAUNCH(true, false);
  else MSCL_FUSED_LAUNCH(false, false);
#undef MSCL_FUSED_LAUNCH
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

// argument checks + Params of one fused job
static int fused_job(Params &p, const float *d_q, const float *d_kpos, int32_t M, const float *d_queue, const int32_t *d_birth,
                     const int64_t *d_qstate, int64_t K_local, float inv_T, float key_norm_bound, const int32_t *d_dup_slot,
                     int32_t dup_age, float *d_ws, float *d_part, int32_t rows_per_group, int32_t with_grad, int32_t flags,
                     float *d_row_loss, float *d_rowaux, float *d_group_out) {
  MSCL_CHECK_ARG(d_q && d_kpos && d_queue && d_birth && d_qstate && d_ws && d_row_loss && d_rowaux && d_group_out,
                 "null pointer");
  MSCL_CHECK_ARG(!with_grad || d_part, "with_grad needs the slab buffer d_part");
  MSCL_CHECK_ARG(M > 0 && K_local > 0, "bad M=%d K_local=%lld", M, (long long)K_local);
  MSCL_CHECK_ARG(K_local < (1ll << 31), "K_local too large for a TMA coordinate");
  MSCL_CHECK_ARG(rows_per_group > 0 && M % rows_per_group == 0, "M=%d must be a multiple of rows_per_group=%d", M,
                 rows_per_group);
  MSCL_CHECK_ARG(inv_T > 0.f && key_norm_bound > 0.f, "bad inv_T / key_norm_bound");
  MSCL_CHECK_ARG((((uintptr_t)d_q | (uintptr_t)d_kpos | (uintptr_t)d_queue | (uintptr_t)d_birth | (uintptr_t)d_ws |
                   (uintptr_t)d_part | (uintptr_t)d_rowaux | (uintptr_t)d_group_out) & 15) == 0,
                 "q/kpos/queue/birth/ws/part/rowaux/group_out must be 16-byte aligned");
  p = Params();
  p.q = d_q;
  p.kpos = d_kpos;
  p.birth = d_birth;
  p.qstate = d_qstate;
  p.dup_slot = d_dup_slot;
  p.ws = d_ws;
  p.part = with_grad ? d_part : nullptr;
  p.row_loss = d_row_loss;
  p.rowaux = d_rowaux;
  p.group_out = d_group_out;
  p.K_local = K_local;
  p.shard_begin = 0;
  p.M = M;
  p.rows_per_group = rows_per_group;
  p.dup_age = dup_age;
  p.flags = flags;
  p.inv_T = inv_T;
  p.key_norm_bound = key_norm_bound;
  p.row_split = 0;
  p.ex_cta = -1;
  p.pre_scale = 1.f;
  return MSCL_OK;
}

}  // namespace tcf
}  // namespace mscl

extern "C" int mscl_infonce_fused(const float *d_q, const float *d_kpos, int32_t M, const float *d_queue,
                                  const int32_t *d_birth, const int64_t *d_qstate, int64_t K_local, float inv_T,
                                  float key_norm_bound, const int32_t *d_dup_slot, int32_t dup_age, float *d_ws,
                                  float *d_part, int32_t n_part, int32_t rows_per_group, int32_t with_grad, int32_t flags,
                                  float *d_row_loss, float *d_rowaux, float *d_group_out, mscl_stream_t stream) {
  using namespace mscl::tcf;
  Params p;
  int rc = fused_job(p, d_q, d_kpos, M, d_queue, d_birth, d_qstate, K_local, inv_T, key_norm_bound, d_dup_slot, dup_age, d_ws,
                     d_part, rows_per_group, with_grad, flags, d_row_loss, d_rowaux, d_group_out);
  if (rc) return rc;
  return launch(true, with_grad != 0, 1, &p, &d_queue, n_part, mscl::as_stream(stream));
}

extern "C" int mscl_infonce_fused_multi_x(int32_t n_jobs, const float *const *d_q, const float *const *d_kpos, const int32_t *M,
                                          const float *const *d_queue, const int32_t *const *d_birth,
                                          const int64_t *const *d_qstate, const int64_t *K_local, const float *inv_T,
                                          const float *key_norm_bound, const int32_t *const *d_dup_slot,
                                          const int32_t *dup_age, float *const *d_ws, float *const *d_part, int32_t n_part,
                                          const int32_t *rows_per_group, int32_t with_grad, const int32_t *flags,
                                          float *const *d_row_loss, float *const *d_rowaux, float *const *d_group_out,
                                          int32_t x_job, const float *d_xkeys, const int32_t *d_xbirth, int64_t rep_begin,
                                          int32_t rep_n, int32_t row_split, mscl_stream_t stream) {
  using namespace mscl::tcf;
  MSCL_CHECK_ARG(x_job >= -1 && x_job < n_jobs, "x_job=%d must be -1 or a job index below %d", x_job, n_jobs);
  MSCL_CHECK_ARG(n_jobs >= 1 && n_jobs <= kMaxJobs, "n_jobs=%d must be in [1, %d]", n_jobs, kMaxJobs);
  MSCL_CHECK_ARG(d_q && d_kpos && M && d_queue && d_birth && d_qstate && K_local && inv_T && key_norm_bound && d_dup_slot &&
                     dup_age && d_ws && d_part && rows_per_group && flags && d_row_loss && d_rowaux && d_group_out,
                 "null table");
  Params ps[kMaxJobs];
  for (int j = 0; j < n_jobs; ++j) {
    int rc = fused_job(ps[j], d_q[j], d_kpos[j], M[j], d_queue[j], d_birth[j], d_qstate[j], K_local[j], inv_T[j],
                       key_norm_bound[j], d_dup_slot[j], dup_age[j], d_ws[j], d_part[j], rows_per_group[j], with_grad, flags[j],
                       d_row_loss[j], d_rowaux[j], d_group_out[j]);
    if (rc) return rc;
    for (int i = 0; i < j; ++i)
      MSCL_CHECK_ARG(d_ws[i] != d_ws[j], "jobs %d and %d share a workspace", i, j);
  }
  if (x_job >= 0) {
    Params &p = ps[x_job];
    MSCL_CHECK_ARG(d_xkeys && d_xbirth && ((uintptr_t)d_xkeys & 15) == 0, "d_xkeys (16-byte aligned) and d_xbirth are required");
    MSCL_CHECK_ARG(rep_n >= 1 && rep_n <= kTile, "rep_n=%d must be in [1, %d]", rep_n, kTile);
    MSCL_CHECK_ARG(rep_begin >= 0 && rep_begin + rep_n <= p.K_local, "replaced slots [%lld, +%d) leave the queue",
                   (long long)rep_begin, rep_n);
    MSCL_CHECK_ARG(row_split >= 0 && row_split <= p.M, "row_split=%d must be in [0, M=%d]", row_split, p.M);
    p.xbirth = d_xbirth;
    p.rep_begin = rep_begin;
    p.rep_n = rep_n;
    p.row_split = row_split;
    p.pre_scale = 1.0f / 0.99999f;
  }
  return launch(true, with_grad != 0, n_jobs, ps, d_queue, n_part, mscl::as_stream(stream), x_job, d_xkeys);
}

extern "C" int mscl_infonce_fused_multi(int32_t n_jobs, const float *const *d_q, const float *const *d_kpos, const int32_t *M,
                                        const float *const *d_queue, const int32_t *const *d_birth,
                                        const int64_t *const *d_qstate, const int64_t *K_local, const float *inv_T,
                                        const float *key_norm_bound, const int32_t *const *d_dup_slot, const int32_t *dup_age,
                                        float *const *d_ws, float *const *d_part, int32_t n_part,
                                        const int32_t *rows_per_group, int32_t with_grad, const int32_t *flags,
                                        float *const *d_row_loss, float *const *d_rowaux, float *const *d_group_out,
                                        mscl_stream_t stream) {
  return mscl_infonce_fused_multi_x(n_jobs, d_q, d_kpos, M, d_queue, d_birth, d_qstate, K_local, inv_T, key_norm_bound,
                                    d_dup_slot, dup_age, d_ws, d_part, n_part, rows_per_group, with_grad, flags, d_row_loss,
                                    d_rowaux, d_group_out, -1, nullptr, nullptr, 0, 0, 0, stream);
}

extern "C" int mscl_infonce_pass(const float *d_qpack, int32_t M, const float *d_queue, const int32_t *d_birth,
                                 const int64_t *d_qstate, int64_t K_local, int64_t shard_begin, float inv_T,
                                 float *d_part, int32_t n_part, int32_t with_grad, int32_t flags, mscl_stream_t stream) {
  using namespace mscl::tcf;
  MSCL_CHECK_ARG(d_qpack && d_queue && d_birth && d_qstate && d_part, "null pointer");
  MSCL_CHECK_ARG(M > 0 && K_local > 0, "bad M=%d K_local=%lld", M, (long long)K_local);
  MSCL_CHECK_ARG(K_local < (1ll << 31), "K_local too large for a TMA coordinate");
  MSCL_CHECK_ARG(inv_T > 0.f, "bad inv_T");
  MSCL_CHECK_ARG((((uintptr_t)d_qpack | (uintptr_t)d_queue | (uintptr_t)d_birth | (uintptr_t)d_part) & 15) == 0,
                 "qpack/queue/birth/part must be 16-byte aligned");
  Params p = {};
  p.q = d_qpack;
  p.birth = d_birth;
  p.qstate = d_qstate;
  p.part = d_part;
  p.K_local = K_local;
  p.shard_begin = shard_begin;
  p.M = M;
  p.rows_per_group = 1;
  p.flags = flags;
  p.inv_T = inv_T;
  p.key_norm_bound = 1.f;
  return launch(false, with_grad != 0, 1, &p, &d_queue, n_part, mscl::as_stream(stream));
}

// How many CTAs along the keys mscl_infonce_fused / mscl_infonce_pass should be given.
extern "C" int mscl_infonce_fused_parts(int32_t M, int64_t K_local, int32_t num_sms) {
  using namespace mscl::tcf;
  if (M <= 0 || K_local <= 0 || num_sms <= 0) return mscl::set_err(MSCL_EINVAL, "bad M / K_local / num_sms");
  const int64_t n_units = (K_local + kUnit - 1) / kUnit;
  const int row_blocks = (M + kRows - 1) / kRows;
  int64_t gx = num_sms / row_blocks;
  if (gx < 1) gx = 1;
  if (gx > n_units) gx = n_units;
  return (int)gx;
}

// The same for several jobs in one launch: the SMs are divided over all row blocks of all jobs.
extern "C" int mscl_infonce_fused_parts_multi(int32_t n_jobs, const int32_t *M, const int64_t *K_local, int32_t num_sms) {
  using namespace mscl::tcf;
  if (n_jobs < 1 || n_jobs > kMaxJobs || !M || !K_local || num_sms <= 0) return mscl::set_err(MSCL_EINVAL, "bad job table / num_sms");
  int64_t rbs = 0, max_units = 0;
  for (int j = 0; j < n_jobs; ++j) {
    if (M[j] <= 0 || K_local[j] <= 0) return mscl::set_err(MSCL_EINVAL, "bad M / K_local of job %d", j);
    rbs += (M[j] + kRows - 1) / kRows;
    const int64_t u = (K_local[j] + kUnit - 1) / kUnit;
    max_units = u > max_units ? u : max_units;
  }
  int64_t gx = num_sms / rbs;
  if (gx < 1) gx = 1;
  if (gx > max_units) gx = max_units;
  return (int)gx;
}

#ifdef MSCL_TC_TIMELINE
extern "C" int mscl_debug_timeline_fused(unsigned long long *host_out, int n) {
  MSCL_CUDA(cudaMemcpyFromSymbol(host_out, mscl::tcf::g_timeline_f, sizeof(unsigned long long) * n));
  return MSCL_OK;
}
#endif
