// K1 (support kernels): prep, CUDA-core validation twin of the tcgen05 pass,
// finalize and backward scaling for the fused InfoNCE objective.
// Contract and formulas: include/mscl_b200.h (K1 block) and DESIGN.md section 4.
#include "common.cuh"

namespace mscl {

constexpr int kC = MSCL_DIM;
constexpr int kLd = MSCL_PACK_LD;

// ---- 1. prep --------------------------------------------------------------
// part A (one warp per query row): qpack row = tf32-rounded q | pos2 | shift2 | dup slot | dup decay
// part B (one thread per key)     : dscale[j]
__global__ void __launch_bounds__(256)
infonce_prep_kernel(const float *__restrict__ q, const float *__restrict__ kpos, int M,
                    const int32_t *__restrict__ birth, const int64_t *__restrict__ qstate,
                    int64_t K_local, float inv_T, float key_norm_bound,
                    float *__restrict__ qpack, float *__restrict__ dscale,
                    const int32_t *__restrict__ dup_slot, int dup_age,
                    float *const *__restrict__ peer_qpack, int n_peers, int row_offset) {
  // Wait first, trigger second: the tcgen05 pass launched behind this grid prefetches queue tiles BEFORE its own
  // dependency wait, so it may only start once everything older than this grid (an enqueue into that queue, say)
  // has completed -- which is what this grid's wait establishes.
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int warps_total = gridDim.x * 8;
  const float sc = inv_T * kLog2e;
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < M; row += warps_total) {
    const float4 a = reinterpret_cast<const float4 *>(q + (int64_t)row * kC)[lane];
    const float4 b = reinterpret_cast<const float4 *>(kpos + (int64_t)row * kC)[lane];
    float d = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    float ss = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    d = warp_sum(d);
    ss = warp_sum(ss);
    float4 r = make_float4(to_tf32_rn(a.x), to_tf32_rn(a.y), to_tf32_rn(a.z), to_tf32_rn(a.w));
    float *dst = qpack + (int64_t)row * kLd;
    reinterpret_cast<float4 *>(dst)[lane] = r;
    // [130] = global queue slot holding a copy of this row's positive key (int bits, -1 = none),
    // [131] = that copy's decay factor 0.99999^age
    const int dup = dup_slot != nullptr ? dup_slot[row] : -1;
    const float4 tail = make_float4(d * sc, sqrtf(ss) * key_norm_bound * sc, __int_as_float(dup),
                                    exp2f((float)dup_age * kLog2Decay));
    if (lane == 0) reinterpret_cast<float4 *>(dst)[32] = tail;
    // sharded queue: the row goes straight into every rank's gathered table over NVLink (peer stores replace
    // the all_gather; a device barrier follows before any rank's pass reads its table)
    for (int p = 0; p < n_peers; ++p) {
      float *pd = peer_qpack[p] + (int64_t)(row_offset + row) * kLd;
      reinterpret_cast<float4 *>(pd)[lane] = r;
      if (lane == 0) reinterpret_cast<float4 *>(pd)[32] = tail;
    }
  }
  if (dscale == nullptr) return;      // the single-launch pass derives the per-key scale from birth[] itself
  const int64_t n_enq = qstate[1];
  const int64_t tid = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * 256;
  // dscale is padded to a multiple of 128 floats (the tcgen05 pass reads whole 128-key
  // slices); the pad must be finite because P' = p * dscale feeds the second GEMM.
  const int64_t K_pad = (K_local + 127) / 128 * 128;
  for (int64_t j = tid; j < K_pad; j += nthreads) {
    float v = 0.f;
    if (j < K_local) {
      const float age = (float)(n_enq - (int64_t)birth[j]);
      v = exp2f(age * kLog2Decay) * sc;
    }
    dscale[j] = v;
  }
}

// ---- 2'. CUDA-core twin of the tcgen05 pass (validation only) ---------------
// One CTA per 128-key tile, 128 threads.  Phase 1: thread <-> key.  Phase 2:
// thread <-> channel.  fp32 FMA everywhere.
constexpr int kSimtKeys = 128;
__global__ void __launch_bounds__(128)
infonce_partial_simt_kernel(const float *__restrict__ qpack, int M,
                            const float *__restrict__ queue,
                            const float *__restrict__ dscale, int64_t K_local,
                            int64_t shard_begin, float *__restrict__ acc, int with_grad) {
  extern __shared__ float sm[];
  float *tile = sm;                       // [128][129]
  float *qrow = tile + kSimtKeys * 129;   // [128]
  float *pbuf = qrow + kC;                // [128]
  float *red = pbuf + kSimtKeys;          // [8]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t k0 = (int64_t)blockIdx.x * kSimtKeys;
  for (int i = tid; i < kSimtKeys * (kC / 4); i += 128) {
    const int r = i / (kC / 4), c4 = i - r * (kC / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k0 + r < K_local) v = __ldg(reinterpret_cast<const float4 *>(queue + (k0 + r) * kC) + c4);
    float *d = tile + r * 129 + c4 * 4;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  const bool valid = (k0 + tid) < K_local;
  const float ds = valid ? dscale[k0 + tid] : 0.f;
  __syncthreads();
  for (int i = 0; i < M; ++i) {
    const float *qp = qpack + (int64_t)i * kLd;
    qrow[tid] = qp[tid];
    const float pos2 = qp[kC], shift2 = qp[kC + 1];
    const int dup = __float_as_int(qp[kC + 2]);
    const bool is_dup = dup >= 0 && (int64_t)dup - shard_begin == k0 + tid;
    __syncthreads();
    float s = 0.f;
    const float *kr = tile + tid * 129;
#pragma unroll 8
    for (int c = 0; c < kC; ++c) s = fmaf(qrow[c], kr[c], s);
    s *= ds;
    float p = (valid && !is_dup) ? exp2f(s - shift2) : 0.f;      // the duplicate is handled exactly by finalize
    float cnt = (valid && !is_dup && s > pos2) ? 1.f : 0.f;
    pbuf[tid] = p * ds;
    float ps = warp_sum(p), cs = warp_sum(cnt);
    if (lane == 0) { red[warp] = ps; red[4 + warp] = cs; }
    __syncthreads();
    if (tid == 0) {
      atomicAdd(acc + (int64_t)i * kLd + kC, red[0] + red[1] + red[2] + red[3]);
      atomicAdd(acc + (int64_t)i * kLd + kC + 1, red[4] + red[5] + red[6] + red[7]);
    }
    if (with_grad) {
      float o = 0.f;
#pragma unroll 8
      for (int j = 0; j < kSimtKeys; ++j) o = fmaf(pbuf[j], tile[j * 129 + tid], o);
      atomicAdd(acc + (int64_t)i * kLd + tid, o);
    }
    __syncthreads();
  }
}

// ---- 2''. sum of the per-CTA partial slabs (sharded path: before the reduce-scatter) ----
__global__ void __launch_bounds__(160)
infonce_reduce_kernel(const float *__restrict__ part, int n_part, int M, float *__restrict__ acc) {
  const int i = blockIdx.x, c = threadIdx.x;
  if (c >= kLd) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const int64_t slab = (int64_t)M * kLd;
  const float *src = part + (int64_t)i * kLd + c;
  int p = 0;
  for (; p + 4 <= n_part; p += 4) {   // fixed order, 4 independent loads in flight
    a0 += src[(p + 0) * slab];
    a1 += src[(p + 1) * slab];
    a2 += src[(p + 2) * slab];
    a3 += src[(p + 3) * slab];
  }
  for (; p < n_part; ++p) a0 += src[p * slab];
  acc[(int64_t)i * kLd + c] = (a0 + a1) + (a2 + a3);
}

// ---- 2x. fused reduce + scatter over NVLink (sharded path) ----------------------------------
// Row i of the gathered query table belongs to rank i / M_local.  This rank's partial result for it (the sum of its
// CTA slabs, fixed order) is stored straight into the OWNER's accumulator, slot `my_rank`:
//   peer_acc[owner][my_rank][i % M_local][0:132]
// so the reduce-scatter is peer stores from the kernel that does the reduction; the owner's finalize then sums
// the G slots in rank order (bit-reproducible, unlike a ring reduce-scatter whose order depends on the rank).
__global__ void __launch_bounds__(160)
infonce_reduce_scatter_kernel(const float *__restrict__ part, int n_part, int M_all, int M_local,
                              float *const *__restrict__ peer_acc, int my_rank) {
  const int i = blockIdx.x, c = threadIdx.x;
  if (c >= kLd) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const int64_t slab = (int64_t)M_all * kLd;
  const float *src = part + (int64_t)i * kLd + c;
  int p = 0;
  for (; p + 4 <= n_part; p += 4) {
    a0 += src[(p + 0) * slab];
    a1 += src[(p + 1) * slab];
    a2 += src[(p + 2) * slab];
    a3 += src[(p + 3) * slab];
  }
  for (; p < n_part; ++p) a0 += src[p * slab];
  const int owner = i / M_local, r = i - owner * M_local;
  float *dst = peer_acc[owner] + ((int64_t)my_rank * M_local + r) * kLd + c;
  *dst = (a0 + a1) + (a2 + a3);
}

// ---- 3. finalize -----------------------------------------------------------
// One CTA per query row, kFinGroups groups of 128 threads (thread <-> channel): group g sums the
// slabs p = g, g+G, g+2G, ... of [n_part][M][132] with all its loads in flight at once (the sum over
// 148 slabs is latency-bound otherwise), the groups are combined through shared memory in a fixed
// order (bit-reproducible), then group 0 computes the row's loss and gradient; the last CTA to
// finish reduces the per-row losses / hit flags into per-group means in row order.
constexpr int kFinGroups = 8;
constexpr int kFinMaxPer = 24;     // slabs per group held in registers per round
__global__ void __launch_bounds__(128 * kFinGroups)
infonce_finalize_kernel(const float *__restrict__ qpack, const float *__restrict__ kpos,
                        float *__restrict__ part, int n_part, int M, int rows_per_group,
                        float inv_T, int with_grad, float *__restrict__ row_loss,
                        float *__restrict__ dq_unit, float *__restrict__ group_out) {
  const int i = blockIdx.x, c = threadIdx.x & 127, g = threadIdx.x >> 7;
  pdl_wait();      // launched early (PDL): the partial slabs are complete and visible from here on
  pdl_trigger();   // a following prep may be launched; it waits for this grid before reading or writing anything
  const float *qp = qpack + (int64_t)i * kLd;
  const int64_t slab = (int64_t)M * kLd;
  const float *src = part + (int64_t)i * kLd;
  __shared__ float s_o[kFinGroups][128];
  __shared__ float s_st[kFinGroups][2];
  {
    float o = 0.f, st = 0.f;
    const bool stat_thread = c < 2;
    for (int p0 = g; p0 < n_part; p0 += kFinGroups * kFinMaxPer) {
      float v[kFinMaxPer], w[kFinMaxPer];
#pragma unroll
      for (int u = 0; u < kFinMaxPer; ++u) {
        const int p = p0 + u * kFinGroups;
        v[u] = (with_grad && p < n_part) ? __ldg(src + p * slab + c) : 0.f;
        w[u] = (stat_thread && p < n_part) ? __ldg(src + p * slab + kC + c) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < kFinMaxPer; ++u) {
        o += v[u];
        st += w[u];
      }
    }
    s_o[g][c] = o;
    if (stat_thread) s_st[g][c] = st;
  }
  __syncthreads();
  if (g != 0) return;
  float o = 0.f, sum = 0.f, cnt = 0.f;
#pragma unroll
  for (int k = 0; k < kFinGroups; ++k) {
    o += s_o[k][c];
    sum += s_st[k][0];
    cnt += s_st[k][1];
  }
  const float pos2 = qp[kC], shift2 = qp[kC + 1];
  // A queue entry that IS this row's positive key (enqueued earlier in the step, moco.py:437)
  // scores pos * 0.99999^age in the reference: above the positive iff pos < 0, and with almost
  // the positive's own probability -- in the gradient the two nearly cancel against the "- 1" of
  // the positive, so a tf32-level error on that one column shows up amplified by 1 / (1 - p_pos - p_dup).
  // The queue pass therefore leaves the column out altogether (no hit test, p = 0), and its exact
  // fp32 contribution is added here: logit = pos * decay (the entry is a bit-exact copy of kpos).
  if (__float_as_int(qp[kC + 2]) >= 0) {
    const float dsc = qp[kC + 3];
    if (pos2 * dsc > pos2) cnt += 1.f;
    const float e_dup = exp2f(fmaf(pos2, dsc, -shift2));
    sum += e_dup;
    if (with_grad) o = fmaf(e_dup * (dsc * inv_T * kLog2e), kpos[(int64_t)i * kC + c], o);
  }
  const float e0 = exp2f(pos2 - shift2);
  const float Z = e0 + sum;
  const float inv_Z = 1.0f / Z;
  const float p0 = e0 * inv_Z;
  // d loss_i / d q = (1/T) [ (p0 - 1) kpos + sum_j p_j decay_j queue_j ];  the
  // accumulated O carries decay_j * log2e / T, hence the ln2 factor.
  const float grad = (p0 - 1.0f) * kpos[(int64_t)i * kC + c] * inv_T + o * inv_Z * kLn2;
  dq_unit[(int64_t)i * kC + c] = grad / (float)rows_per_group;
  __shared__ int is_last;
  if (c == 0) {
    const float loss = (shift2 + log2f(Z) - pos2) * kLn2;
    // row_loss[i] = loss, row_loss[M + i] = #negatives above the positive
    row_loss[i] = loss;
    row_loss[M + i] = cnt;
    __threadfence();
    unsigned *counter = reinterpret_cast<unsigned *>(part + kC + 2);   // row 0 of slab 0, a spare (zero) column
    const unsigned done = atomicAdd(counter, 1u);
    is_last = (done == (unsigned)M - 1u);
    if (is_last) *counter = 0u;
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");   // group 0 only (the other groups have exited)
  if (!is_last) return;
  __threadfence();
  const int n_groups = M / rows_per_group;
  // one warp per loss term: lane l sums rows l, l+32, ... in order, then a fixed shuffle tree (bit-reproducible)
  const int lane = c & 31;
  for (int gidx = c >> 5; gidx < n_groups; gidx += 4) {
    float sl = 0.f, s1 = 0.f, s5 = 0.f;
    for (int r = lane; r < rows_per_group; r += 32) {
      const int row = gidx * rows_per_group + r;
      sl += __ldcg(row_loss + row);
      const float k = __ldcg(row_loss + M + row);
      s1 += (k < 1.f) ? 1.f : 0.f;
      s5 += (k < 5.f) ? 1.f : 0.f;
    }
    sl = warp_sum(sl);
    s1 = warp_sum(s1);
    s5 = warp_sum(s5);
    if (lane == 0) {
      const float inv = 1.0f / (float)rows_per_group;
      *reinterpret_cast<float4 *>(group_out + gidx * 4) = make_float4(sl * inv, s1 * inv, s5 * inv, 0.f);
    }
  }
}

__global__ void __launch_bounds__(128)
infonce_bwd_kernel(const float *__restrict__ dq_unit, const float *__restrict__ gout, int M,
                   int rows_per_group, float *__restrict__ dq) {
  const int i = blockIdx.x, c = threadIdx.x;
  dq[(int64_t)i * kC + c] = dq_unit[(int64_t)i * kC + c] * __ldg(gout + i / rows_per_group);
}

// Backward of the single-launch form: dq_i = gout[group(i)] * (ck_i kpos_i + co_i sum_p part[p][i][0:128]).
// One CTA per query row, kFinGroups groups of 128 threads (thread <-> channel): group g sums the slabs p = g, g+G, ...
// with all its loads in flight at once, the groups are combined through shared memory in a fixed order
// (bit-reproducible given the slabs).
__global__ void __launch_bounds__(128 * kFinGroups)
infonce_bwd_slabs_kernel(const float *__restrict__ part, int n_part, int M, const float *__restrict__ kpos,
                         const float *__restrict__ rowaux, const float *__restrict__ gout, int rows_per_group,
                         float *__restrict__ dq) {
  const int i = blockIdx.x, c = threadIdx.x & 127, g = threadIdx.x >> 7;
  const int64_t slab = (int64_t)M * kLd;
  const float *src = part + (int64_t)i * kLd;
  __shared__ float s_o[kFinGroups][128];
  float o = 0.f;
  pdl_wait();      // may have been launched early behind the forward kernel (back-to-back in benchmarks)
  pdl_trigger();
  for (int p0 = g; p0 < n_part; p0 += kFinGroups * kFinMaxPer) {
    float v[kFinMaxPer];
#pragma unroll
    for (int u = 0; u < kFinMaxPer; ++u) {
      const int p = p0 + u * kFinGroups;
      v[u] = p < n_part ? __ldg(src + p * slab + c) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kFinMaxPer; ++u) o += v[u];
  }
  s_o[g][c] = o;
  __syncthreads();
  if (g != 0) return;
  o = 0.f;
#pragma unroll
  for (int k = 0; k < kFinGroups; ++k) o += s_o[k][c];
  const float4 ra = __ldg(reinterpret_cast<const float4 *>(rowaux) + i);
  dq[(int64_t)i * kC + c] = __ldg(gout + i / rows_per_group) * fmaf(ra.z, kpos[(int64_t)i * kC + c], ra.w * o);
}

// The same for the jobs of one mscl_infonce_fused_multi(_x) launch, in ONE launch: blockIdx.x runs over the rows of all jobs.
struct BwdJobs {
  const float *part[4];
  const float *kpos[4];
  const float *rowaux[4];
  const float *gout[4];
  float *dq[4];
  int M[4];
  int rows_per_group[4];
  int row_begin[5];
  int n_jobs;
  int n_part;
};

__global__ void __launch_bounds__(128 * kFinGroups)
infonce_bwd_slabs_multi_kernel(const __grid_constant__ BwdJobs jb) {
  int j = 0;
  while (j + 1 < jb.n_jobs && (int)blockIdx.x >= jb.row_begin[j + 1]) ++j;
  const int i = (int)blockIdx.x - jb.row_begin[j], c = threadIdx.x & 127, g = threadIdx.x >> 7;
  const int M = jb.M[j], n_part = jb.n_part;
  const int64_t slab = (int64_t)M * kLd;
  const float *src = jb.part[j] + (int64_t)i * kLd;
  __shared__ float s_o[kFinGroups][128];
  float o = 0.f;
  pdl_wait();
  pdl_trigger();
  for (int p0 = g; p0 < n_part; p0 += kFinGroups * kFinMaxPer) {
    float v[kFinMaxPer];
#pragma unroll
    for (int u = 0; u < kFinMaxPer; ++u) {
      const int p = p0 + u * kFinGroups;
      v[u] = p < n_part ? __ldg(src + p * slab + c) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kFinMaxPer; ++u) o += v[u];
  }
  s_o[g][c] = o;
  __syncthreads();
  if (g != 0) return;
  o = 0.f;
#pragma unroll
  for (int k = 0; k < kFinGroups; ++k) o += s_o[k][c];
  const float4 ra = __ldg(reinterpret_cast<const float4 *>(jb.rowaux[j]) + i);
  jb.dq[j][(int64_t)i * kC + c] =
      __ldg(jb.gout[j] + i / jb.rows_per_group[j]) * fmaf(ra.z, jb.kpos[j][(int64_t)i * kC + c], ra.w * o);
}

}  // namespace mscl

extern "C" {

int mscl_infonce_bwd_slabs_multi(int32_t n_jobs, const float *const *d_part, int32_t n_part, const int32_t *M,
                                 const float *const *d_kpos, const float *const *d_rowaux, const float *const *d_gout,
                                 const int32_t *rows_per_group, float *const *d_dq, mscl_stream_t stream) {
  MSCL_CHECK_ARG(n_jobs >= 1 && n_jobs <= 4, "n_jobs=%d must be in [1, 4]", n_jobs);
  MSCL_CHECK_ARG(d_part && M && d_kpos && d_rowaux && d_gout && rows_per_group && d_dq, "null table");
  MSCL_CHECK_ARG(n_part > 0, "n_part=%d must be positive", n_part);
  mscl::BwdJobs jb = {};
  jb.n_jobs = n_jobs;
  jb.n_part = n_part;
  for (int j = 0; j < n_jobs; ++j) {
    MSCL_CHECK_ARG(d_part[j] && d_kpos[j] && d_rowaux[j] && d_gout[j] && d_dq[j], "null pointer in job %d", j);
    MSCL_CHECK_ARG(M[j] > 0 && rows_per_group[j] > 0 && M[j] % rows_per_group[j] == 0,
                   "job %d: M=%d must be a multiple of rows_per_group=%d", j, M[j], rows_per_group[j]);
    MSCL_CHECK_ARG(((uintptr_t)d_rowaux[j] & 15) == 0, "rowaux must be 16-byte aligned");
    jb.part[j] = d_part[j];
    jb.kpos[j] = d_kpos[j];
    jb.rowaux[j] = d_rowaux[j];
    jb.gout[j] = d_gout[j];
    jb.dq[j] = d_dq[j];
    jb.M[j] = M[j];
    jb.rows_per_group[j] = rows_per_group[j];
    jb.row_begin[j + 1] = jb.row_begin[j] + M[j];
  }
  MSCL_CUDA(mscl::launch_pdl(mscl::infonce_bwd_slabs_multi_kernel, dim3((unsigned)jb.row_begin[n_jobs]),
                             dim3(128 * mscl::kFinGroups), 0, mscl::as_stream(stream), jb));
  return MSCL_OK;
}

int mscl_infonce_bwd_slabs(const float *d_part, int32_t n_part, int32_t M, const float *d_kpos, const float *d_rowaux,
                           const float *d_gout, int32_t rows_per_group, float *d_dq, mscl_stream_t stream) {
  MSCL_CHECK_ARG(d_part && d_kpos && d_rowaux && d_gout && d_dq, "null pointer");
  MSCL_CHECK_ARG(M > 0 && n_part > 0 && rows_per_group > 0 && M % rows_per_group == 0,
                 "M=%d must be a multiple of rows_per_group=%d (n_part=%d)", M, rows_per_group, n_part);
  MSCL_CHECK_ARG(((uintptr_t)d_rowaux & 15) == 0, "rowaux must be 16-byte aligned");
  MSCL_CUDA(mscl::launch_pdl(mscl::infonce_bwd_slabs_kernel, dim3(M), dim3(128 * mscl::kFinGroups), 0, mscl::as_stream(stream),
                             d_part, n_part, M, d_kpos, d_rowaux, d_gout, rows_per_group, d_dq));
  return MSCL_OK;
}

int mscl_infonce_prep(const float *d_q, const float *d_kpos, int32_t M,
                      const int32_t *d_birth, const int64_t *d_qstate, int64_t K_local,
                      float inv_T, float key_norm_bound, float *d_qpack, float *d_dscale,
                      const int32_t *d_dup_slot, int32_t dup_age, float *const *d_peer_qpack,
                      int32_t n_peers, int32_t row_offset, mscl_stream_t stream) {
  MSCL_CHECK_ARG(d_q && d_kpos && d_birth && d_qstate && d_qpack, "null pointer");
  MSCL_CHECK_ARG(M > 0 && K_local > 0, "bad M=%d K_local=%lld", M, (long long)K_local);
  MSCL_CHECK_ARG(inv_T > 0.f && key_norm_bound > 0.f, "bad inv_T / key_norm_bound");
  MSCL_CHECK_ARG((n_peers == 0) || (d_peer_qpack != nullptr && n_peers > 0 && row_offset >= 0), "bad peer table");
  MSCL_CHECK_ARG((((uintptr_t)d_q | (uintptr_t)d_kpos | (uintptr_t)d_qpack) & 15) == 0,
                 "q/kpos/qpack must be 16-byte aligned");
  int64_t want = d_dscale ? (K_local + 255) / 256 : 1;
  if ((M + 7) / 8 > want) want = (M + 7) / 8;
  if (want > 1184) want = 1184;
  MSCL_CUDA(mscl::launch_pdl(mscl::infonce_prep_kernel, dim3((unsigned)want), dim3(256), 0, mscl::as_stream(stream), d_q,
                             d_kpos, M, d_birth, d_qstate, K_local, inv_T, key_norm_bound, d_qpack, d_dscale, d_dup_slot,
                             dup_age, d_peer_qpack, n_peers, row_offset));
  return MSCL_OK;
}

int mscl_infonce_partial_simt(const float *d_qpack, int32_t M, const float *d_queue,
                              const float *d_dscale, int64_t K_local, int64_t shard_begin,
                              float *d_acc, int32_t with_grad, mscl_stream_t stream) {
  MSCL_CHECK_ARG(d_qpack && d_queue && d_dscale && d_acc, "null pointer");
  MSCL_CHECK_ARG(M > 0 && K_local > 0, "bad M=%d K_local=%lld", M, (long long)K_local);
  const size_t smem = sizeof(float) * (mscl::kSimtKeys * 129 + mscl::kC + mscl::kSimtKeys + 8);
  MSCL_CUDA(mscl::ensure_dyn_smem(mscl::infonce_partial_simt_kernel, smem));
  const int64_t blocks = (K_local + mscl::kSimtKeys - 1) / mscl::kSimtKeys;
  mscl::infonce_partial_simt_kernel<<<(unsigned)blocks, 128, smem, mscl::as_stream(stream)>>>(
      d_qpack, M, d_queue, d_dscale, K_local, shard_begin, d_acc, with_grad);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

int mscl_infonce_reduce(const float *d_part, int32_t n_part, int32_t M, float *d_acc,
                        mscl_stream_t stream) {
  MSCL_CHECK_ARG(d_part && d_acc, "null pointer");
  MSCL_CHECK_ARG(M > 0 && n_part > 0, "bad M=%d n_part=%d", M, n_part);
  mscl::infonce_reduce_kernel<<<M, 160, 0, mscl::as_stream(stream)>>>(d_part, n_part, M, d_acc);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

int mscl_infonce_reduce_scatter(const float *d_part, int32_t n_part, int32_t M_all, int32_t M_local,
                                float *const *d_peer_acc, int32_t my_rank, mscl_stream_t stream) {
  MSCL_CHECK_ARG(d_part && d_peer_acc, "null pointer");
  MSCL_CHECK_ARG(M_all > 0 && n_part > 0 && M_local > 0 && M_all % M_local == 0 && my_rank >= 0 && my_rank < M_all / M_local,
                 "bad M_all=%d M_local=%d n_part=%d rank=%d", M_all, M_local, n_part, my_rank);
  mscl::infonce_reduce_scatter_kernel<<<M_all, 160, 0, mscl::as_stream(stream)>>>(d_part, n_part, M_all, M_local, d_peer_acc,
                                                                                  my_rank);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

int mscl_infonce_finalize(const float *d_qpack, const float *d_kpos, float *d_part,
                          int32_t n_part, int32_t M, int32_t rows_per_group, float inv_T,
                          int32_t with_grad, float *d_row_loss, float *d_dq_unit,
                          float *d_group_out, mscl_stream_t stream) {
  MSCL_CHECK_ARG(d_qpack && d_kpos && d_part && d_row_loss && d_dq_unit && d_group_out,
                 "null pointer");
  MSCL_CHECK_ARG(M > 0 && n_part > 0 && rows_per_group > 0 && M % rows_per_group == 0,
                 "M=%d must be a multiple of rows_per_group=%d (n_part=%d)", M, rows_per_group, n_part);
  MSCL_CUDA(mscl::launch_pdl(mscl::infonce_finalize_kernel, dim3(M), dim3(128 * mscl::kFinGroups), 0,
                             mscl::as_stream(stream), d_qpack, d_kpos, d_part, n_part, M, rows_per_group, inv_T,
                             with_grad, d_row_loss, d_dq_unit, d_group_out));
  return MSCL_OK;
}

int mscl_infonce_bwd(const float *d_dq_unit, const float *d_gout, int32_t M,
                     int32_t rows_per_group, float *d_dq, mscl_stream_t stream) {
  MSCL_CHECK_ARG(d_dq_unit && d_gout && d_dq, "null pointer");
  MSCL_CHECK_ARG(M > 0 && rows_per_group > 0 && M % rows_per_group == 0,
                 "M=%d must be a multiple of rows_per_group=%d", M, rows_per_group);
  mscl::infonce_bwd_kernel<<<M, 128, 0, mscl::as_stream(stream)>>>(d_dq_unit, d_gout, M,
                                                                  rows_per_group, d_dq);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

}  // extern "C"
