// K3: FRA -- flow rotation augmentation on the GPU
// (mmaction/datasets/pipelines/transforms_motion.py:103-142 and norm_flow :7-29).
//
// Element-wise, HBM-bound: 128-bit loads of u and v, 128-bit stores of the four
// output planes.  Per (u,v) pixel pair: 8 B read in the max pre-pass, 8 B read +
// 16 B written in the apply pass.  Arithmetic follows the reference's float32
// sequence (separate multiplies and subtract/add, IEEE divide), so the rotated and
// normalised values agree with NumPy float32 to the last bit when the same
// cos/sin float32 constants are used.
#include "common.cuh"

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
