// K3: FRA -- flow rotation augmentation on the GPU
// (mmaction/datasets/pipelines/transforms_motion.py:103-142 and norm_flow :7-29).
//
// Element-wise, HBM-bound: 128-bit loads of u and v, 128-bit stores of the four
// output planes.  Per (u,v) pixel pair: 8 B read in the max pre-pass, 8 B read +
// 16 B written in the apply pass.  Arithmetic follows the reference's float32
// sequence (separate multiplies and subtract/add, IEEE divide), so the rotated and
// normalised values agree with NumPy float32 to the last bit when the same
// cos/sin float32 constants are used.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace mscl {

struct UV4 {
  float4 u, v;
};

// Load 4 consecutive pixels of frame (n, t) starting at pixel p.
template <int LAYOUT>
__device__ __forceinline__ UV4 load_uv4(const float *__restrict__ flow, int n, int t, int T,
                                        int HW, int p) {
  UV4 r;
  if (LAYOUT == 0) {
    const float *pu = flow + (((int64_t)n * 2 + 0) * T + t) * HW + p;
    const float *pv = flow + (((int64_t)n * 2 + 1) * T + t) * HW + p;
    r.u = ldg_stream(reinterpret_cast<const float4 *>(pu));
    r.v = ldg_stream(reinterpret_cast<const float4 *>(pv));
  } else {
    const float *pp = flow + (((int64_t)n * T + t) * HW + p) * 2;
    const float4 a = ldg_stream(reinterpret_cast<const float4 *>(pp));
    const float4 b = ldg_stream(reinterpret_cast<const float4 *>(pp) + 1);
    r.u = make_float4(a.x, a.z, b.x, b.z);
    r.v = make_float4(a.y, a.w, b.y, b.w);
  }
  return r;
}

__device__ __forceinline__ float rot_u(float u, float v, float c, float s) {
  return __fsub_rn(__fmul_rn(c, u), __fmul_rn(s, v));
}
__device__ __forceinline__ float rot_v(float u, float v, float c, float s) {
  return __fadd_rn(__fmul_rn(s, u), __fmul_rn(c, v));
}
__device__ __forceinline__ float radius(float u, float v) {
  return __fsqrt_rn(__fadd_rn(__fmul_rn(u, u), __fmul_rn(v, v)));
}

// grid: (blocks per frame, N*T).  maxrad[frame][0] = max |base|, [1] = max |rotated|.
template <int LAYOUT>
__global__ void __launch_bounds__(256)
fra_maxrad_kernel(const float *__restrict__ flow, const int32_t *__restrict__ cid,
                  const float *__restrict__ cs, float *__restrict__ maxrad, int T, int HW) {
  const int frame = blockIdx.y;
  const int n = frame / T, t = frame - n * T;
  const int id = cid[n];
  const float c = cs[2 * id], s = cs[2 * id + 1];
  float mb = 0.f, mr = 0.f;
  for (int p = (blockIdx.x * 256 + threadIdx.x) * 4; p < HW; p += gridDim.x * 1024) {
    const UV4 x = load_uv4<LAYOUT>(flow, n, t, T, HW, p);
    const float us[4] = {x.u.x, x.u.y, x.u.z, x.u.w};
    const float vs[4] = {x.v.x, x.v.y, x.v.z, x.v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      mb = fmaxf(mb, radius(us[e], vs[e]));
      mr = fmaxf(mr, radius(rot_u(us[e], vs[e], c, s), rot_v(us[e], vs[e], c, s)));
    }
  }
  mb = warp_max(mb);
  mr = warp_max(mr);
  __shared__ float sb[8], sr[8];
  if ((threadIdx.x & 31) == 0) {
    sb[threadIdx.x >> 5] = mb;
    sr[threadIdx.x >> 5] = mr;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      mb = fmaxf(mb, sb[w]);
      mr = fmaxf(mr, sr[w]);
    }
    // radii are >= 0: float order == signed-int order of the bit patterns
    atomicMax(reinterpret_cast<int *>(maxrad + 2 * frame), __float_as_int(mb));
    atomicMax(reinterpret_cast<int *>(maxrad + 2 * frame + 1), __float_as_int(mr));
  }
}

__device__ __forceinline__ float4 div4(float4 a, float d) {
  return make_float4(__fdiv_rn(a.x, d), __fdiv_rn(a.y, d), __fdiv_rn(a.z, d), __fdiv_rn(a.w, d));
}

template <int LAYOUT>
__global__ void __launch_bounds__(256)
fra_apply_kernel(const float *__restrict__ flow, const int32_t *__restrict__ cid,
                 const float *__restrict__ cs, const float *__restrict__ maxrad,
                 float *__restrict__ out, int T, int HW) {
  const int frame = blockIdx.y;
  const int n = frame / T, t = frame - n * T;
  const int id = cid[n];
  const float c = cs[2 * id], s = cs[2 * id + 1];
  const float db = __fadd_rn(maxrad[2 * frame], 1e-5f);
  const float dr = __fadd_rn(maxrad[2 * frame + 1], 1e-5f);
  const int T2 = 2 * T;
  float *ou = out + (((int64_t)n * 2 + 0) * T2) * HW;
  float *ov = out + (((int64_t)n * 2 + 1) * T2) * HW;
  for (int p = (blockIdx.x * 256 + threadIdx.x) * 4; p < HW; p += gridDim.x * 1024) {
    const UV4 x = load_uv4<LAYOUT>(flow, n, t, T, HW, p);
    float4 ru, rv;
    ru.x = rot_u(x.u.x, x.v.x, c, s); rv.x = rot_v(x.u.x, x.v.x, c, s);
    ru.y = rot_u(x.u.y, x.v.y, c, s); rv.y = rot_v(x.u.y, x.v.y, c, s);
    ru.z = rot_u(x.u.z, x.v.z, c, s); rv.z = rot_v(x.u.z, x.v.z, c, s);
    ru.w = rot_u(x.u.w, x.v.w, c, s); rv.w = rot_v(x.u.w, x.v.w, c, s);
    stg_stream(reinterpret_cast<float4 *>(ou + (int64_t)t * HW + p), div4(x.u, db));
    stg_stream(reinterpret_cast<float4 *>(ov + (int64_t)t * HW + p), div4(x.v, db));
    stg_stream(reinterpret_cast<float4 *>(ou + (int64_t)(T + t) * HW + p), div4(ru, dr));
    stg_stream(reinterpret_cast<float4 *>(ov + (int64_t)(T + t) * HW + p), div4(rv, dr));
  }
}

// ---- one-pass FRA: a thread-block cluster per frame ------------------------------------------
// The two-kernel form reads every (u,v) twice (max pre-pass, then apply): 32 B per pixel pair.
// Here the 8 CTAs of a cluster each keep one eighth of the frame in shared memory, exchange their
// partial maxima through distributed shared memory, and write the four output planes from the staged
// copy: 8 B read + 16 B written = 24 B per pixel pair, one launch, no scratch, no memset.
// max sqrt(x) == sqrt(max x) exactly (IEEE sqrt is monotonic), so only the squared radii are compared.
constexpr int kFraMaxCluster = 8;
template <int LAYOUT>
__global__ void __launch_bounds__(1024)
fra_fused_kernel(const float *__restrict__ flow, const int32_t *__restrict__ cid,
                 const float *__restrict__ cs, float *__restrict__ out, int T, int HW, int chunk) {
  extern __shared__ float4 stage[];          // u4[chunk/4] | v4[chunk/4]
  __shared__ float cl_max[2];                // this CTA's (base, rotated) squared maxima: read by the whole cluster
  __shared__ float sb[32], sr[32];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();       // the cluster spans grid.x
  const int nranks = (int)cluster.num_blocks();
  const int nthr = blockDim.x;
  const int frame = blockIdx.y;
  const int n = frame / T, t = frame - n * T;
  const int id = cid[n];
  const float c = cs[2 * id], s = cs[2 * id + 1];
  const int p0 = rank * chunk;
  const int p1 = min(p0 + chunk, HW);
  // stage this CTA's share of the frame with bulk async copies (one elected thread, one mbarrier): the whole
  // share is in flight at once and never passes through registers
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
  const uint32_t npx = (uint32_t)(p1 > p0 ? p1 - p0 : 0);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(npx * 8u) : "memory");
    if (npx) {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(stage);
      if (LAYOUT == 0) {
        const float *pu = flow + (((int64_t)n * 2 + 0) * T + t) * HW + p0;
        const float *pv = flow + (((int64_t)n * 2 + 1) * T + t) * HW + p0;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "l"(pu), "r"(npx * 4u), "r"(bar_a) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         dst + (uint32_t)chunk * 4u), "l"(pv), "r"(npx * 4u), "r"(bar_a) : "memory");
      } else {
        const float *pp = flow + (((int64_t)n * T + t) * HW + p0) * 2;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "l"(pp), "r"(npx * 8u), "r"(bar_a) : "memory");
      }
    }
  }
  __syncthreads();
  {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar_a) : "memory");
  }
  auto staged = [&](int g) {       // 4 consecutive pixels of the staged share
    UV4 x;
    if (LAYOUT == 0) {
      x.u = stage[g];
      x.v = stage[(chunk >> 2) + g];
    } else {
      const float4 a = stage[2 * g], b2 = stage[2 * g + 1];
      x.u = make_float4(a.x, a.z, b2.x, b2.z);
      x.v = make_float4(a.y, a.w, b2.y, b2.w);
    }
    return x;
  };
  float mb = 0.f, mr = 0.f;
  for (int g = threadIdx.x; p0 + 4 * g < p1; g += nthr) {
    const UV4 x = staged(g);
    const float us[4] = {x.u.x, x.u.y, x.u.z, x.u.w};
    const float vs[4] = {x.v.x, x.v.y, x.v.z, x.v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      mb = fmaxf(mb, __fadd_rn(__fmul_rn(us[e], us[e]), __fmul_rn(vs[e], vs[e])));
      const float ru = rot_u(us[e], vs[e], c, s), rv = rot_v(us[e], vs[e], c, s);
      mr = fmaxf(mr, __fadd_rn(__fmul_rn(ru, ru), __fmul_rn(rv, rv)));
    }
  }
  mb = warp_max(mb);
  mr = warp_max(mr);
  if ((threadIdx.x & 31) == 0) {
    sb[threadIdx.x >> 5] = mb;
    sr[threadIdx.x >> 5] = mr;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = nthr >> 5;
    mb = threadIdx.x < nw ? sb[threadIdx.x] : 0.f;
    mr = threadIdx.x < nw ? sr[threadIdx.x] : 0.f;
    mb = warp_max(mb);
    mr = warp_max(mr);
    if (threadIdx.x == 0) {
      cl_max[0] = mb;
      cl_max[1] = mr;
    }
  }
  float fb = 0.f, fr = 0.f;
  if (nranks > 1) {
    cluster.sync();
    for (int r = 0; r < nranks; ++r) {
      const float *remote = cluster.map_shared_rank(cl_max, r);
      fb = fmaxf(fb, remote[0]);
      fr = fmaxf(fr, remote[1]);
    }
    cluster.sync();                          // nobody leaves (or reuses cl_max) while a peer may still read it
  } else {
    __syncthreads();
    fb = cl_max[0];
    fr = cl_max[1];
  }
  const float db = __fadd_rn(__fsqrt_rn(fb), 1e-5f);
  const float dr = __fadd_rn(__fsqrt_rn(fr), 1e-5f);
  const int T2 = 2 * T;
  float *ou = out + (((int64_t)n * 2 + 0) * T2) * HW;
  float *ov = out + (((int64_t)n * 2 + 1) * T2) * HW;
  for (int g = threadIdx.x; p0 + 4 * g < p1; g += nthr) {
    const int p = p0 + 4 * g;
    const UV4 x = staged(g);
    const float4 u = x.u, v = x.v;
    float4 ru, rv;
    ru.x = rot_u(u.x, v.x, c, s); rv.x = rot_v(u.x, v.x, c, s);
    ru.y = rot_u(u.y, v.y, c, s); rv.y = rot_v(u.y, v.y, c, s);
    ru.z = rot_u(u.z, v.z, c, s); rv.z = rot_v(u.z, v.z, c, s);
    ru.w = rot_u(u.w, v.w, c, s); rv.w = rot_v(u.w, v.w, c, s);
    stg_stream(reinterpret_cast<float4 *>(ou + (int64_t)t * HW + p), div4(u, db));
    stg_stream(reinterpret_cast<float4 *>(ov + (int64_t)t * HW + p), div4(v, db));
    stg_stream(reinterpret_cast<float4 *>(ou + (int64_t)(T + t) * HW + p), div4(ru, dr));
    stg_stream(reinterpret_cast<float4 *>(ov + (int64_t)(T + t) * HW + p), div4(rv, dr));
  }
}

// rotation only, planar in -> planar out (16 B per pixel pair)
__global__ void __launch_bounds__(256)
fra_rotate_kernel(const float *__restrict__ flow, const int32_t *__restrict__ cid,
                  const float *__restrict__ cs, float *__restrict__ out, int T, int HW) {
  const int frame = blockIdx.y;
  const int n = frame / T, t = frame - n * T;
  const int id = cid[n];
  const float c = cs[2 * id], s = cs[2 * id + 1];
  const int64_t offu = (((int64_t)n * 2 + 0) * T + t) * HW;
  const int64_t offv = (((int64_t)n * 2 + 1) * T + t) * HW;
  for (int p = (blockIdx.x * 256 + threadIdx.x) * 4; p < HW; p += gridDim.x * 1024) {
    const UV4 x = load_uv4<0>(flow, n, t, T, HW, p);
    float4 ru, rv;
    ru.x = rot_u(x.u.x, x.v.x, c, s); rv.x = rot_v(x.u.x, x.v.x, c, s);
    ru.y = rot_u(x.u.y, x.v.y, c, s); rv.y = rot_v(x.u.y, x.v.y, c, s);
    ru.z = rot_u(x.u.z, x.v.z, c, s); rv.z = rot_v(x.u.z, x.v.z, c, s);
    ru.w = rot_u(x.u.w, x.v.w, c, s); rv.w = rot_v(x.u.w, x.v.w, c, s);
    stg_stream(reinterpret_cast<float4 *>(out + offu + p), ru);
    stg_stream(reinterpret_cast<float4 *>(out + offv + p), rv);
  }
}

static int frame_blocks(int HW) {
  int b = (HW / 4 + 255) / 256;
  return b < 1 ? 1 : b;
}

}  // namespace mscl

extern "C" {

static int fra_check(const void *flow, const void *cid, const void *cs, int N, int T, int HW,
                     int layout) {
  MSCL_CHECK_ARG(flow && cid && cs, "null pointer");
  MSCL_CHECK_ARG(N > 0 && T > 0 && HW > 0, "bad N=%d T=%d HW=%d", N, T, HW);
  MSCL_CHECK_ARG(HW % 4 == 0, "H*W=%d must be a multiple of 4 (128-bit accesses)", HW);
  MSCL_CHECK_ARG(layout == 0 || layout == 1, "layout=%d", layout);
  MSCL_CHECK_ARG(((uintptr_t)flow % 16) == 0, "flow must be 16-byte aligned");
  MSCL_CHECK_ARG((int64_t)N * T <= 65535, "N*T=%lld exceeds grid.y", (long long)N * T);
  return MSCL_OK;
}

int mscl_fra_maxrad(const float *d_flow, const int32_t *d_cid, const float *d_cs,
                    float *d_maxrad, int32_t N, int32_t T, int32_t HW, int32_t layout,
                    mscl_stream_t stream) {
  int rc = fra_check(d_flow, d_cid, d_cs, N, T, HW, layout);
  if (rc) return rc;
  MSCL_CHECK_ARG(d_maxrad, "null maxrad");
  cudaStream_t s = mscl::as_stream(stream);
  MSCL_CUDA(cudaMemsetAsync(d_maxrad, 0, sizeof(float) * 2 * (size_t)N * T, s));
  dim3 grid(mscl::frame_blocks(HW), N * T);
  if (layout == 0)
    mscl::fra_maxrad_kernel<0><<<grid, 256, 0, s>>>(d_flow, d_cid, d_cs, d_maxrad, T, HW);
  else
    mscl::fra_maxrad_kernel<1><<<grid, 256, 0, s>>>(d_flow, d_cid, d_cs, d_maxrad, T, HW);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

int mscl_fra_apply(const float *d_flow, const int32_t *d_cid, const float *d_cs,
                   const float *d_maxrad, float *d_out, int32_t N, int32_t T, int32_t HW,
                   int32_t layout, mscl_stream_t stream) {
  int rc = fra_check(d_flow, d_cid, d_cs, N, T, HW, layout);
  if (rc) return rc;
  MSCL_CHECK_ARG(d_maxrad && d_out, "null pointer");
  MSCL_CHECK_ARG(((uintptr_t)d_out % 16) == 0, "out must be 16-byte aligned");
  dim3 grid(mscl::frame_blocks(HW), N * T);
  cudaStream_t s = mscl::as_stream(stream);
  if (layout == 0)
    mscl::fra_apply_kernel<0><<<grid, 256, 0, s>>>(d_flow, d_cid, d_cs, d_maxrad, d_out, T, HW);
  else
    mscl::fra_apply_kernel<1><<<grid, 256, 0, s>>>(d_flow, d_cid, d_cs, d_maxrad, d_out, T, HW);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

int mscl_fra_fused(const float *d_flow, const int32_t *d_cid, const float *d_cs, float *d_out,
                   int32_t N, int32_t T, int32_t HW, int32_t layout, mscl_stream_t stream) {
  int rc = fra_check(d_flow, d_cid, d_cs, N, T, HW, layout);
  if (rc) return rc;
  MSCL_CHECK_ARG(d_out && ((uintptr_t)d_out % 16) == 0, "out must be non-null, 16-byte aligned");
  // cluster size: the smallest power of two whose per-CTA share of the frame fits 110 KB (two CTAs per SM)
  static int env_cluster = -1, env_threads = -1;      // tuning overrides
  if (env_cluster < 0) {
    const char *e = getenv("MSCL_FRA_CLUSTER");
    env_cluster = e ? atoi(e) : 0;
    e = getenv("MSCL_FRA_THREADS");
    env_threads = e ? atoi(e) : 0;
  }
  int cluster = 1;
  while (cluster < mscl::kFraMaxCluster && (size_t)((HW / 4 + cluster - 1) / cluster) * 32 > 110 * 1024) cluster *= 2;
  if (env_cluster > 0) cluster = env_cluster;
  const int chunk = ((HW / 4 + cluster - 1) / cluster) * 4;      // pixels per CTA
  const size_t smem = (size_t)chunk * 8;
  MSCL_CHECK_ARG(smem <= 200 * 1024, "frame of %d pixels does not fit a cluster's shared memory: use maxrad + apply", HW);
  int threads = chunk / 4 >= 512 ? 512 : 256;
  if (env_threads > 0) threads = env_threads;
  if (layout == 0)
    MSCL_CUDA(mscl::ensure_dyn_smem(mscl::fra_fused_kernel<0>, smem));
  else
    MSCL_CUDA(mscl::ensure_dyn_smem(mscl::fra_fused_kernel<1>, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cluster, N * T);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = mscl::as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (layout == 0)
    MSCL_CUDA(cudaLaunchKernelEx(&cfg, mscl::fra_fused_kernel<0>, d_flow, d_cid, d_cs, d_out, (int)T, (int)HW, chunk));
  else
    MSCL_CUDA(cudaLaunchKernelEx(&cfg, mscl::fra_fused_kernel<1>, d_flow, d_cid, d_cs, d_out, (int)T, (int)HW, chunk));
  return MSCL_OK;
}

int mscl_fra_rotate(const float *d_flow, const int32_t *d_cid, const float *d_cs,
                    float *d_out, int32_t N, int32_t T, int32_t HW, mscl_stream_t stream) {
  int rc = fra_check(d_flow, d_cid, d_cs, N, T, HW, 0);
  if (rc) return rc;
  MSCL_CHECK_ARG(d_out && ((uintptr_t)d_out % 16) == 0, "out must be non-null, 16-byte aligned");
  dim3 grid(mscl::frame_blocks(HW), N * T);
  mscl::fra_rotate_kernel<<<grid, 256, 0, mscl::as_stream(stream)>>>(d_flow, d_cid, d_cs,
                                                                     d_out, T, HW);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

}  // extern "C"
