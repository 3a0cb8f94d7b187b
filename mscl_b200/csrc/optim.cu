// K10: gradient-norm clipping + SGD-with-momentum step as two multi-tensor launches
// (the reference trains with torch.optim.SGD(lr=0.02, momentum=0.9, weight_decay=1e-4) and
//  optimizer_config.grad_clip = dict(max_norm=40, norm_type=2): configs/recognition/moco/mscl_r18_cosm_lr2e-2.py:112-119,
//  applied by mmcv's OptimizerHook -> torch.nn.utils.clip_grad_norm_).  SURVEY.md section 8f-3.
//
//   pass 1  grad_sqnorm_multi : per-CTA partial sums of g^2 (fixed order -> bit-reproducible), 4 B/element
//   pass 2  clip_sgd_multi    : coef = min(1, max_norm / (||g|| + 1e-6));  g <- g*coef;  d = g + wd*p;
//                               buf <- momentum*buf + d  (buf <- d on the first step);  p <- p - lr*buf
//                               12 B read + 12 B written per element.
// PyTorch's foreach implementation makes ~10 passes over the same 150 MB (norms, scale, add, mul, add, add).
// Same chunk tables as the EMA kernel (one CTA per 16 Ki-element chunk of any tensor).
#include "common.cuh"

namespace mscl {

__global__ void __launch_bounds__(256)
grad_sqnorm_multi_kernel(const float *const *__restrict__ g_ptrs, const int64_t *__restrict__ sizes,
                         const int32_t *__restrict__ blk_tensor, const int64_t *__restrict__ blk_start, int chunk_elems,
                         float *__restrict__ partial) {
  const int t = blk_tensor[blockIdx.x];
  const int64_t start = blk_start[blockIdx.x];
  const float *__restrict__ g = g_ptrs[t] + start;
  int64_t n = sizes[t] - start;
  if (n > chunk_elems) n = chunk_elems;
  float acc = 0.f;
  if ((((uintptr_t)g) & 15) == 0) {
    const int64_t nv = n >> 2;
    const float4 *g4 = reinterpret_cast<const float4 *>(g);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int64_t v = threadIdx.x;
    for (; v + 3 * 256 < nv; v += 4 * 256) {
      const float4 x0 = g4[v], x1 = g4[v + 256], x2 = g4[v + 512], x3 = g4[v + 768];
      a0 += (x0.x * x0.x + x0.y * x0.y) + (x0.z * x0.z + x0.w * x0.w);
      a1 += (x1.x * x1.x + x1.y * x1.y) + (x1.z * x1.z + x1.w * x1.w);
      a2 += (x2.x * x2.x + x2.y * x2.y) + (x2.z * x2.z + x2.w * x2.w);
      a3 += (x3.x * x3.x + x3.y * x3.y) + (x3.z * x3.z + x3.w * x3.w);
    }
    for (; v < nv; v += 256) {
      const float4 x0 = g4[v];
      a0 += (x0.x * x0.x + x0.y * x0.y) + (x0.z * x0.z + x0.w * x0.w);
    }
    acc = (a0 + a1) + (a2 + a3);
    for (int64_t i = (nv << 2) + threadIdx.x; i < n; i += 256) acc += g[i] * g[i];
  } else {
    for (int64_t i = threadIdx.x; i < n; i += 256) acc += g[i] * g[i];
  }
  acc = warp_sum(acc);
  __shared__ float s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tsum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tsum += s[w];
    partial[blockIdx.x] = tsum;
  }
}

// stats[0] = total L2 norm of the gradients, stats[1] = clip coefficient (<= 1)
__global__ void __launch_bounds__(1024)
grad_norm_finish_kernel(const float *__restrict__ partial, int n, float max_norm, float *__restrict__ stats) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += 1024) acc += partial[i];     // fixed order per thread, fixed tree below
  acc = warp_sum(acc);
  __shared__ float s[32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = s[threadIdx.x];
    v = warp_sum(v);
    if (threadIdx.x == 0) {
      const float norm = sqrtf(v);
      const float coef = max_norm / (norm + 1e-6f);
      stats[0] = norm;
      stats[1] = coef < 1.f ? coef : 1.f;
    }
  }
}

__device__ __forceinline__ void sgd1(float &g, float &p, float &b, float coef, float wd, float mom, float lr, bool first) {
  g = g * coef;
  const float d = g + wd * p;
  b = first ? d : mom * b + d;
  p = p - lr * b;
}

__global__ void __launch_bounds__(256)
clip_sgd_multi_kernel(float *const *__restrict__ g_ptrs, float *const *__restrict__ p_ptrs, float *const *__restrict__ b_ptrs,
                      const int64_t *__restrict__ sizes, const int32_t *__restrict__ blk_tensor,
                      const int64_t *__restrict__ blk_start, int chunk_elems, const float *__restrict__ stats, float wd,
                      float mom, float lr, int first) {
  const int t = blk_tensor[blockIdx.x];
  const int64_t start = blk_start[blockIdx.x];
  float *__restrict__ g = g_ptrs[t] + start;
  float *__restrict__ p = p_ptrs[t] + start;
  float *__restrict__ b = b_ptrs[t] + start;
  int64_t n = sizes[t] - start;
  if (n > chunk_elems) n = chunk_elems;
  const float coef = stats != nullptr ? stats[1] : 1.f;
  const bool f = first != 0;
  if ((((uintptr_t)g | (uintptr_t)p | (uintptr_t)b) & 15) == 0) {
    const int64_t nv = n >> 2;
    float4 *g4 = reinterpret_cast<float4 *>(g), *p4 = reinterpret_cast<float4 *>(p), *b4 = reinterpret_cast<float4 *>(b);
    for (int64_t v = threadIdx.x; v < nv; v += 2 * 256) {
      float4 gg[2], pp[2], bb[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t i = v + u * 256;
        if (i < nv) {
          gg[u] = g4[i];
          pp[u] = p4[i];
          bb[u] = f ? make_float4(0.f, 0.f, 0.f, 0.f) : b4[i];
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t i = v + u * 256;
        if (i < nv) {
          sgd1(gg[u].x, pp[u].x, bb[u].x, coef, wd, mom, lr, f);
          sgd1(gg[u].y, pp[u].y, bb[u].y, coef, wd, mom, lr, f);
          sgd1(gg[u].z, pp[u].z, bb[u].z, coef, wd, mom, lr, f);
          sgd1(gg[u].w, pp[u].w, bb[u].w, coef, wd, mom, lr, f);
          g4[i] = gg[u];
          p4[i] = pp[u];
          b4[i] = bb[u];
        }
      }
    }
    for (int64_t i = (nv << 2) + threadIdx.x; i < n; i += 256) {
      float bi = f ? 0.f : b[i];
      sgd1(g[i], p[i], bi, coef, wd, mom, lr, f);
      b[i] = bi;
    }
  } else {
    for (int64_t i = threadIdx.x; i < n; i += 256) {
      float bi = f ? 0.f : b[i];
      sgd1(g[i], p[i], bi, coef, wd, mom, lr, f);
      b[i] = bi;
    }
  }
}

}  // namespace mscl

extern "C" {

int mscl_grad_norm_multi(const float *const *d_g_ptrs, const int64_t *d_sizes, const int32_t *d_blk_tensor,
                         const int64_t *d_blk_start, int32_t n_blocks, int32_t chunk_elems, float max_norm,
                         float *d_partial, float *d_stats, mscl_stream_t stream) {
  MSCL_CHECK_ARG(d_g_ptrs && d_sizes && d_blk_tensor && d_blk_start && d_partial && d_stats, "null pointer");
  MSCL_CHECK_ARG(n_blocks > 0 && chunk_elems > 0 && chunk_elems % 4 == 0 && max_norm > 0.f, "bad n_blocks / chunk_elems / max_norm");
  cudaStream_t s = mscl::as_stream(stream);
  mscl::grad_sqnorm_multi_kernel<<<n_blocks, 256, 0, s>>>(d_g_ptrs, d_sizes, d_blk_tensor, d_blk_start, chunk_elems, d_partial);
  MSCL_LAUNCH_CHECK();
  mscl::grad_norm_finish_kernel<<<1, 1024, 0, s>>>(d_partial, n_blocks, max_norm, d_stats);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

int mscl_clip_sgd_multi(float *const *d_g_ptrs, float *const *d_p_ptrs, float *const *d_buf_ptrs, const int64_t *d_sizes,
                        const int32_t *d_blk_tensor, const int64_t *d_blk_start, int32_t n_blocks, int32_t chunk_elems,
                        const float *d_stats, float weight_decay, float momentum, float lr, int32_t first_step,
                        mscl_stream_t stream) {
  MSCL_CHECK_ARG(d_g_ptrs && d_p_ptrs && d_buf_ptrs && d_sizes && d_blk_tensor && d_blk_start, "null pointer");
  MSCL_CHECK_ARG(n_blocks > 0 && chunk_elems > 0 && chunk_elems % 4 == 0, "bad n_blocks / chunk_elems");
  mscl::clip_sgd_multi_kernel<<<n_blocks, 256, 0, mscl::as_stream(stream)>>>(d_g_ptrs, d_p_ptrs, d_buf_ptrs, d_sizes, d_blk_tensor,
                                                                            d_blk_start, chunk_elems, d_stats, weight_decay,
                                                                            momentum, lr, first_step);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

}  // extern "C"
