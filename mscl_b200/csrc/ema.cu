// K4: multi-tensor momentum EMA of the key encoder (moco.py:408-421).
//
// One launch covers every (param_k, param_q) pair.  A CTA owns one chunk of one
// tensor (table lookup), streams it with 128-bit loads/stores, 4 independent
// vectors in flight per thread.  12 B/element of HBM traffic, nothing else.
// Arithmetic mirrors the reference exactly: two fp32 multiplies, one fp32 add,
// no FMA contraction -> bit-identical to `k*m + q*(1-m)` in PyTorch.
#include "common.cuh"

namespace mscl {

__device__ __forceinline__ float ema1(float k, float q, float m, float om) {
  return __fadd_rn(__fmul_rn(k, m), __fmul_rn(q, om));
}

__global__ void __launch_bounds__(256)
ema_multi_kernel(float *const *__restrict__ k_ptrs, const float *const *__restrict__ q_ptrs,
                 const int64_t *__restrict__ sizes, const int32_t *__restrict__ blk_tensor,
                 const int64_t *__restrict__ blk_start, int chunk_elems, float m, float om) {
  const int t = blk_tensor[blockIdx.x];
  const int64_t start = blk_start[blockIdx.x];
  float *__restrict__ k = k_ptrs[t] + start;
  const float *__restrict__ q = q_ptrs[t] + start;
  int64_t n = sizes[t] - start;
  if (n > chunk_elems) n = chunk_elems;

  const bool aligned = (((uintptr_t)k | (uintptr_t)q) & 15) == 0;
  if (aligned) {
    const int64_t nv = n >> 2;
    float4 *k4 = reinterpret_cast<float4 *>(k);
    const float4 *q4 = reinterpret_cast<const float4 *>(q);
    for (int64_t v = threadIdx.x; v < nv; v += 4 * 256) {
      float4 a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t i = v + u * 256;
        if (i < nv) {
          a[u] = k4[i];
          b[u] = ldg_stream(q4 + i);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t i = v + u * 256;
        if (i < nv) {
          float4 r;
          r.x = ema1(a[u].x, b[u].x, m, om);
          r.y = ema1(a[u].y, b[u].y, m, om);
          r.z = ema1(a[u].z, b[u].z, m, om);
          r.w = ema1(a[u].w, b[u].w, m, om);
          k4[i] = r;
        }
      }
    }
    for (int64_t i = (nv << 2) + threadIdx.x; i < n; i += 256) k[i] = ema1(k[i], q[i], m, om);
  } else {
    for (int64_t i = threadIdx.x; i < n; i += 256) k[i] = ema1(k[i], q[i], m, om);
  }
}

}  // namespace mscl

extern "C" int mscl_ema_multi(float *const *d_k_ptrs, const float *const *d_q_ptrs,
                              const int64_t *d_sizes, const int32_t *d_blk_tensor,
                              const int64_t *d_blk_start, int32_t n_blocks,
                              int32_t chunk_elems, float m, float one_minus_m,
                              mscl_stream_t stream) {
  MSCL_CHECK_ARG(d_k_ptrs && d_q_ptrs && d_sizes && d_blk_tensor && d_blk_start,
                 "null table pointer");
  MSCL_CHECK_ARG(n_blocks > 0, "n_blocks=%d", n_blocks);
  MSCL_CHECK_ARG(chunk_elems > 0 && chunk_elems % 4 == 0, "chunk_elems=%d must be a positive multiple of 4",
                 chunk_elems);
  mscl::ema_multi_kernel<<<n_blocks, 256, 0, mscl::as_stream(stream)>>>(
      d_k_ptrs, d_q_ptrs, d_sizes, d_blk_tensor, d_blk_start, chunk_elems, m, one_minus_m);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}
