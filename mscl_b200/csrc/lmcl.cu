// K2: LMCL frame-level RGB-vs-flow contrast
// (mmaction/models/heads/local_cl_head.py:57-73 forward, :41-55 loss).
//
//  hw_mean_fwd/bwd : the AdaptiveAvgPool3d((None,1,1)) over H*W.  This is where the
//                    bytes are (51 MB for the r18 config's q_mlvl[0]); one warp per
//                    (n,c,t) row, 128-bit streaming loads, HBM-bound.
//  lmcl            : per clip, on the pooled (C,t) / (C,2t) matrices: L2-normalise
//                    over C, t x 2t similarities / T, cross-entropy against the
//                    diagonal, top-1/5, and the full backward to the pooled
//                    features -- all in one launch, one CTA per clip, everything in
//                    shared memory.  Launch-latency bound (KFLOPs of work).
#include "common.cuh"

namespace mscl {

__global__ void __launch_bounds__(256)
hw_mean_fwd_kernel(const float *__restrict__ x, float *__restrict__ out, int64_t R, int HW,
                   float inv_hw) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= R) return;
  const float *p = x + row * HW;
  float acc = 0.f;
  if ((HW & 3) == 0 && (((uintptr_t)p) & 15) == 0) {
    const float4 *p4 = reinterpret_cast<const float4 *>(p);
    const int nv = HW >> 2;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int v = lane;
    for (; v + 96 < nv; v += 128) {
      const float4 f0 = ldg_stream(p4 + v), f1 = ldg_stream(p4 + v + 32),
                   f2 = ldg_stream(p4 + v + 64), f3 = ldg_stream(p4 + v + 96);
      a0 += (f0.x + f0.y) + (f0.z + f0.w);
      a1 += (f1.x + f1.y) + (f1.z + f1.w);
      a2 += (f2.x + f2.y) + (f2.z + f2.w);
      a3 += (f3.x + f3.y) + (f3.z + f3.w);
    }
    for (; v < nv; v += 32) {
      const float4 f0 = ldg_stream(p4 + v);
      a0 += (f0.x + f0.y) + (f0.z + f0.w);
    }
    acc = (a0 + a1) + (a2 + a3);
  } else {
    for (int i = lane; i < HW; i += 32) acc += __ldg(p + i);
  }
  acc = warp_sum(acc);
  if (lane == 0) out[row] = acc * inv_hw;
}

__global__ void __launch_bounds__(256)
hw_mean_bwd_kernel(const float *__restrict__ gout, float *__restrict__ gx, int64_t R, int HW,
                   float inv_hw) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= R) return;
  const float g = __ldg(gout + row) * inv_hw;
  float *p = gx + row * HW;
  if ((HW & 3) == 0 && (((uintptr_t)p) & 15) == 0) {
    const float4 g4 = make_float4(g, g, g, g);
    float4 *p4 = reinterpret_cast<float4 *>(p);
    for (int v = lane; v < (HW >> 2); v += 32) stg_stream(p4 + v, g4);
  } else {
    for (int i = lane; i < HW; i += 32) p[i] = g;
  }
}

// Small planes (H*W < 128, e.g. the 7x7 flow maps): a warp per 49-element row leaves most lanes idle and
// issues 196-byte requests.  Here a CTA stages 64 consecutive rows -- one contiguous 16-byte-aligned span --
// through shared memory with 128-bit loads, then 4 threads per row sum it (odd H*W: conflict-free pitch).
constexpr int kSmallRows = 64;
__global__ void __launch_bounds__(256)
hw_mean_fwd_small_kernel(const float *__restrict__ x, float *__restrict__ out, int64_t R, int HW,
                         float inv_hw) {
  extern __shared__ float4 buf4[];
  float *buf = reinterpret_cast<float *>(buf4);
  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * kSmallRows;
  const int nrows = (int)((R - row0) < kSmallRows ? (R - row0) : kSmallRows);
  const int nfl = nrows * HW;
  const float *base = x + row0 * HW;
  const int n4 = nfl >> 2;
  for (int v = tid; v < n4; v += 256) buf4[v] = ldg_stream(reinterpret_cast<const float4 *>(base) + v);
  for (int i = (n4 << 2) + tid; i < nfl; i += 256) buf[i] = __ldg(base + i);
  __syncthreads();
  const int r = tid >> 2, part = tid & 3;
  float acc = 0.f;
  if (r < nrows) {
    const float *p = buf + r * HW;
    for (int i = part; i < HW; i += 4) acc += p[i];
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  if (part == 0 && r < nrows) out[row0 + r] = acc * inv_hw;
}

// backward of the same: plain element-wise fill, one float4 per thread, row index by division
__global__ void __launch_bounds__(256)
hw_mean_bwd_small_kernel(const float *__restrict__ gout, float *__restrict__ gx, int64_t total4, int HW,
                         float inv_hw) {
  const int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (v >= total4) return;
  const int64_t e = v << 2;
  const int64_t r = e / HW;
  const int rem = (int)(e - r * HW);
  const float g0 = __ldg(gout + r) * inv_hw;
  float g1 = g0;
  if (rem + 3 >= HW) g1 = __ldg(gout + r + 1) * inv_hw;     // HW >= 4: at most one row boundary inside a float4
  float4 o;
  o.x = g0;
  o.y = (rem + 1 >= HW) ? g1 : g0;
  o.z = (rem + 2 >= HW) ? g1 : g0;
  o.w = (rem + 3 >= HW) ? g1 : g0;
  stg_stream(reinterpret_cast<float4 *>(gx) + v, o);
}

// ---- channels-last (NDHWC) forms: x is [N, T, H*W, C] in memory (torch.channels_last_3d), the pooled output stays
// (N, C, T) row-major like the row-major kernels'.  One CTA per (n, t) plane and 32-channel group: 8 lanes (float4) span
// the 128-byte channel slice of one pixel, 32 pixel groups walk H*W with 4 independent accumulators, a shared-memory
// tree combines them in a fixed order.  The encoders run channels-last, so this reads the feature map where it lies:
// the row-major kernels needed a 2 x 51 MB layout copy in front (forward) and behind (backward) at the r18 sizes.
constexpr int kClGroups = 32;     // pixel groups per CTA (256 threads = 32 groups x 8 lanes)
__global__ void __launch_bounds__(256)
hw_mean_ndhwc_fwd_kernel(const float4 *__restrict__ x, float *__restrict__ out, int C, int T, int HW, float inv_hw) {
  __shared__ float4 red[kClGroups][8];
  const int lane8 = threadIdx.x & 7, grp = threadIdx.x >> 3;
  const int c4_per_row = C >> 2;
  const int cg = blockIdx.x;                       // 32-channel group
  const int64_t plane = blockIdx.y;                // n * T + t
  const float4 *p = x + plane * HW * c4_per_row + cg * 8 + lane8;
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
  int px = grp;
  for (; px + kClGroups < HW; px += 2 * kClGroups) {
    const float4 f0 = ldg_stream(p + (int64_t)px * c4_per_row), f1 = ldg_stream(p + (int64_t)(px + kClGroups) * c4_per_row);
    a0.x += f0.x; a0.y += f0.y; a0.z += f0.z; a0.w += f0.w;
    a1.x += f1.x; a1.y += f1.y; a1.z += f1.z; a1.w += f1.w;
  }
  if (px < HW) {
    const float4 f0 = ldg_stream(p + (int64_t)px * c4_per_row);
    a0.x += f0.x; a0.y += f0.y; a0.z += f0.z; a0.w += f0.w;
  }
  red[grp][lane8] = make_float4(a0.x + a1.x, a0.y + a1.y, a0.z + a1.z, a0.w + a1.w);
  __syncthreads();
#pragma unroll
  for (int half = kClGroups / 2; half > 0; half >>= 1) {
    if (grp < half) {
      const float4 o = red[grp + half][lane8];
      float4 &m = red[grp][lane8];
      m.x += o.x; m.y += o.y; m.z += o.z; m.w += o.w;
    }
    __syncthreads();
  }
  if (grp == 0) {
    const float4 m = red[0][lane8];
    const int64_t n = plane / T;
    const int t = (int)(plane - n * T);
    float *o = out + (n * C + cg * 32 + lane8 * 4) * T + t;      // (N, C, T): channel stride T
    o[0] = m.x * inv_hw;
    o[T] = m.y * inv_hw;
    o[2 * T] = m.z * inv_hw;
    o[3 * T] = m.w * inv_hw;
  }
}

__global__ void __launch_bounds__(256)
hw_mean_ndhwc_bwd_kernel(const float *__restrict__ gout, float4 *__restrict__ gx, int C, int T, int HW, float inv_hw) {
  const int lane8 = threadIdx.x & 7, grp = threadIdx.x >> 3;
  const int c4_per_row = C >> 2;
  const int cg = blockIdx.x;
  const int64_t plane = blockIdx.y;
  const int64_t n = plane / T;
  const int t = (int)(plane - n * T);
  const float *g = gout + (n * C + cg * 32 + lane8 * 4) * T + t;
  const float4 v = make_float4(__ldg(g) * inv_hw, __ldg(g + T) * inv_hw, __ldg(g + 2 * T) * inv_hw, __ldg(g + 3 * T) * inv_hw);
  float4 *p = gx + plane * HW * c4_per_row + cg * 8 + lane8;
  for (int px = grp; px < HW; px += kClGroups) stg_stream(p + (int64_t)px * c4_per_row, v);
}

// One CTA (256 threads) per clip.
// smem floats: y[C*ld] | dy[C*ld] | sim[t*sp] | nrm[tt] | dot[tt] | red[32*3]   with tt = t + t2,
// ld = tt | 1 and sp = t2 | 1: odd pitches, so the column walks (norms, dy) hit 32 different banks.
// Columns 0..t-1 of y are the RGB frames, t..tt-1 the flow frames (base then FRA).
__global__ void __launch_bounds__(1024)
lmcl_kernel(const float *__restrict__ xq, const float *__restrict__ xf, int N, int C, int t,
            int t2, float inv_T, float *__restrict__ out, float *__restrict__ gxq,
            float *__restrict__ gxf, float *__restrict__ part) {
  extern __shared__ float sm[];
  const int tt = t + t2;
  const int ld = tt | 1, sp = t2 | 1;
  float *y = sm;
  float *dy = y + C * ld;
  float *sim = dy + C * ld;
  float *nrm = sim + t * sp;
  float *dot = nrm + tt;
  float *red = dot + tt;
  const int n = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthr = blockDim.x, nwarps = nthr >> 5;
  const float *xq_n = xq + (int64_t)n * C * t;
  const float *xf_n = xf + (int64_t)n * C * t2;

  // 1. stage: y[c*ld + col]
  for (int i = tid; i < C * t; i += nthr) {
    const int c = i / t, col = i - c * t;
    y[c * ld + col] = __ldg(xq_n + i);
  }
  for (int i = tid; i < C * t2; i += nthr) {
    const int c = i / t2, col = i - c * t2;
    y[c * ld + t + col] = __ldg(xf_n + i);
  }
  __syncthreads();
  // 2. column norms (F.normalize: x / max(||x||_2, 1e-12))
  for (int col = warp; col < tt; col += nwarps) {
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float v = y[c * ld + col];
      ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    if (lane == 0) nrm[col] = fmaxf(sqrtf(ss), 1e-12f);
  }
  __syncthreads();
  for (int i = tid; i < C * tt; i += nthr) {
    const int c = i / tt, col = i - c * tt;
    y[c * ld + col] = __fdiv_rn(y[c * ld + col], nrm[col]);
  }
  __syncthreads();
  // 3. similarities / T
  for (int ij = tid; ij < t * t2; ij += nthr) {
    const int i = ij / t2, j = ij - i * t2;
    float a0 = 0.f, a1 = 0.f;
    int c = 0;
    for (; c + 1 < C; c += 2) {
      a0 = fmaf(y[c * ld + i], y[c * ld + t + j], a0);
      a1 = fmaf(y[(c + 1) * ld + i], y[(c + 1) * ld + t + j], a1);
    }
    if (c < C) a0 = fmaf(y[c * ld + i], y[c * ld + t + j], a0);
    sim[i * sp + j] = (a0 + a1) * inv_T;
  }
  __syncthreads();
  // 4. per-row cross-entropy against column i; sim <- d loss / d sim (unit upstream)
  float l_loss = 0.f, l_t1 = 0.f, l_t5 = 0.f;
  const float gscale = inv_T / (float)(N * t);
  for (int i = warp; i < t; i += nwarps) {
    const float spos = sim[i * sp + i];
    float mx = -INFINITY;
    for (int j = lane; j < t2; j += 32) mx = fmaxf(mx, sim[i * sp + j]);
    mx = warp_max(mx);
    float se = 0.f, cnt = 0.f;
    for (int j = lane; j < t2; j += 32) {
      const float s = sim[i * sp + j];
      se += __expf(s - mx);
      cnt += (j != i && s > spos) ? 1.f : 0.f;
    }
    se = warp_sum(se);
    cnt = warp_sum(cnt);
    const float lse = mx + __logf(se);
    for (int j = lane; j < t2; j += 32) {
      const float p = __expf(sim[i * sp + j] - lse);
      sim[i * sp + j] = (p - (j == i ? 1.f : 0.f)) * gscale;
    }
    if (lane == 0) {
      l_loss += lse - spos;
      l_t1 += (cnt < 1.f) ? 1.f : 0.f;
      l_t5 += (cnt < 5.f) ? 1.f : 0.f;
    }
  }
  if (lane == 0) {
    red[warp * 3 + 0] = l_loss;
    red[warp * 3 + 1] = l_t1;
    red[warp * 3 + 2] = l_t5;
  }
  __syncthreads();
  // 5. dy = d loss / d y
  for (int i = tid; i < C * tt; i += nthr) {
    const int c = i / tt, col = i - c * tt;
    float acc = 0.f;
    if (col < t) {
      for (int j = 0; j < t2; ++j) acc = fmaf(sim[col * sp + j], y[c * ld + t + j], acc);
    } else {
      const int j = col - t;
      for (int r = 0; r < t; ++r) acc = fmaf(sim[r * sp + j], y[c * ld + r], acc);
    }
    dy[c * ld + col] = acc;
  }
  __syncthreads();
  // 6. back through the normalisation: dx = (dy - y (y . dy)) / ||x||
  for (int col = warp; col < tt; col += nwarps) {
    float d = 0.f;
    for (int c = lane; c < C; c += 32) d = fmaf(y[c * ld + col], dy[c * ld + col], d);
    d = warp_sum(d);
    if (lane == 0) dot[col] = d;
  }
  __syncthreads();
  float *gq_n = gxq + (int64_t)n * C * t;
  float *gf_n = gxf + (int64_t)n * C * t2;
  for (int i = tid; i < C * t; i += nthr) {        // coalesced stores, one output tensor at a time
    const int c = i / t, col = i - c * t;
    gq_n[i] = (dy[c * ld + col] - y[c * ld + col] * dot[col]) / nrm[col];
  }
  for (int i = tid; i < C * t2; i += nthr) {
    const int c = i / t2, col = t + i - c * t2;
    gf_n[i] = (dy[c * ld + col] - y[c * ld + col] * dot[col]) / nrm[col];
  }
  // 7. per-clip partials; the last CTA to finish sums them with one warp: lane k takes clips k, k+32, ...
  //    in order, then a fixed shuffle tree -> the same bits on every run (no float atomics)
  __shared__ int is_last;
  if (tid == 0) {
    float a = 0.f, b = 0.f, c5 = 0.f;
    for (int w = 0; w < nwarps; ++w) {
      a += red[w * 3 + 0];
      b += red[w * 3 + 1];
      c5 += red[w * 3 + 2];
    }
    *reinterpret_cast<float4 *>(part + n * 4) = make_float4(a, b, c5, 0.f);
    __threadfence();
    unsigned *counter = reinterpret_cast<unsigned *>(part + (int64_t)N * 4);
    const unsigned done = atomicAdd(counter, 1u);
    is_last = (done == (unsigned)N - 1u);
  }
  __syncthreads();
  if (is_last && warp == 0) {
    __threadfence();
    float sl = 0.f, s1 = 0.f, s5 = 0.f;
    for (int k = lane; k < N; k += 32) {
      const float4 v = __ldcg(reinterpret_cast<const float4 *>(part) + k);
      sl += v.x;
      s1 += v.y;
      s5 += v.z;
    }
    sl = warp_sum(sl);
    s1 = warp_sum(s1);
    s5 = warp_sum(s5);
    if (lane == 0) {
      const float inv = 1.f / (float)(N * t);
      out[0] = sl * inv;
      out[1] = s1 * inv;
      out[2] = s5 * inv;
      out[3] = 0.f;
      *reinterpret_cast<unsigned *>(part + (int64_t)N * 4) = 0u;
    }
  }
}

}  // namespace mscl

extern "C" {

int mscl_hw_mean_fwd(const float *d_x, float *d_out, int64_t R, int32_t HW,
                     mscl_stream_t stream) {
  MSCL_CHECK_ARG(d_x && d_out, "null pointer");
  MSCL_CHECK_ARG(R > 0 && HW > 0, "bad R=%lld HW=%d", (long long)R, HW);
  if (HW >= 4 && HW < 128 && (((uintptr_t)d_x) & 15) == 0) {
    const int64_t nb = (R + mscl::kSmallRows - 1) / mscl::kSmallRows;
    MSCL_CHECK_ARG(nb < (1ll << 31), "too many rows");
    mscl::hw_mean_fwd_small_kernel<<<(unsigned)nb, 256, (size_t)mscl::kSmallRows * HW * 4, mscl::as_stream(stream)>>>(
        d_x, d_out, R, HW, 1.0f / (float)HW);
    MSCL_LAUNCH_CHECK();
    return MSCL_OK;
  }
  const int64_t blocks = (R + 7) / 8;
  MSCL_CHECK_ARG(blocks < (1ll << 31), "too many rows");
  mscl::hw_mean_fwd_kernel<<<(unsigned)blocks, 256, 0, mscl::as_stream(stream)>>>(
      d_x, d_out, R, HW, 1.0f / (float)HW);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

int mscl_hw_mean_bwd(const float *d_gout, float *d_gx, int64_t R, int32_t HW,
                     mscl_stream_t stream) {
  MSCL_CHECK_ARG(d_gout && d_gx, "null pointer");
  MSCL_CHECK_ARG(R > 0 && HW > 0, "bad R=%lld HW=%d", (long long)R, HW);
  if (HW >= 4 && HW < 128 && (((uintptr_t)d_gx) & 15) == 0 && ((R * HW) & 3) == 0) {
    const int64_t total4 = (R * HW) >> 2;
    const int64_t nb = (total4 + 255) / 256;
    MSCL_CHECK_ARG(nb < (1ll << 31), "too many elements");
    mscl::hw_mean_bwd_small_kernel<<<(unsigned)nb, 256, 0, mscl::as_stream(stream)>>>(d_gout, d_gx, total4, HW,
                                                                                     1.0f / (float)HW);
    MSCL_LAUNCH_CHECK();
    return MSCL_OK;
  }
  const int64_t blocks = (R + 7) / 8;
  MSCL_CHECK_ARG(blocks < (1ll << 31), "too many rows");
  mscl::hw_mean_bwd_kernel<<<(unsigned)blocks, 256, 0, mscl::as_stream(stream)>>>(
      d_gout, d_gx, R, HW, 1.0f / (float)HW);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

int mscl_hw_mean_ndhwc_fwd(const float *d_x, float *d_out, int64_t N, int32_t C, int32_t T, int32_t HW, mscl_stream_t stream) {
  MSCL_CHECK_ARG(d_x && d_out, "null pointer");
  MSCL_CHECK_ARG(N > 0 && T > 0 && HW > 0 && C > 0 && C % 32 == 0, "bad N=%lld C=%d (multiple of 32) T=%d HW=%d", (long long)N, C, T, HW);
  MSCL_CHECK_ARG(N * T <= 65535, "N*T=%lld exceeds the grid's y extent", (long long)(N * T));
  MSCL_CHECK_ARG((((uintptr_t)d_x) & 15) == 0, "x must be 16-byte aligned");
  mscl::hw_mean_ndhwc_fwd_kernel<<<dim3(C / 32, (unsigned)(N * T)), 256, 0, mscl::as_stream(stream)>>>(
      reinterpret_cast<const float4 *>(d_x), d_out, C, T, HW, 1.0f / (float)HW);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

int mscl_hw_mean_ndhwc_bwd(const float *d_gout, float *d_gx, int64_t N, int32_t C, int32_t T, int32_t HW, mscl_stream_t stream) {
  MSCL_CHECK_ARG(d_gout && d_gx, "null pointer");
  MSCL_CHECK_ARG(N > 0 && T > 0 && HW > 0 && C > 0 && C % 32 == 0, "bad N=%lld C=%d (multiple of 32) T=%d HW=%d", (long long)N, C, T, HW);
  MSCL_CHECK_ARG(N * T <= 65535, "N*T=%lld exceeds the grid's y extent", (long long)(N * T));
  MSCL_CHECK_ARG((((uintptr_t)d_gx) & 15) == 0, "gx must be 16-byte aligned");
  mscl::hw_mean_ndhwc_bwd_kernel<<<dim3(C / 32, (unsigned)(N * T)), 256, 0, mscl::as_stream(stream)>>>(
      d_gout, reinterpret_cast<float4 *>(d_gx), C, T, HW, 1.0f / (float)HW);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

int mscl_lmcl(const float *d_xq, const float *d_xf, int32_t N, int32_t C, int32_t t,
              int32_t t2, float inv_T, float *d_out, float *d_gxq, float *d_gxf,
              float *d_part, mscl_stream_t stream) {
  MSCL_CHECK_ARG(d_xq && d_xf && d_out && d_gxq && d_gxf && d_part, "null pointer");
  MSCL_CHECK_ARG(N > 0 && C > 0 && t > 0 && t2 >= t, "bad N=%d C=%d t=%d t2=%d", N, C, t, t2);
  const int tt = t + t2;
  const size_t smem = sizeof(float) * ((size_t)2 * C * (tt | 1) + (size_t)t * (t2 | 1) + 2 * tt + 96);
  MSCL_CHECK_ARG(smem <= 200 * 1024, "C*(t+t2) too large for one CTA (%zu B of shared memory)",
                 smem);
  MSCL_CUDA(mscl::ensure_dyn_smem(mscl::lmcl_kernel, smem));
  const int threads = (C * tt >= 4096) ? 1024 : ((C * tt >= 2048) ? 512 : 256);     // one thread per ~6 staged elements
  mscl::lmcl_kernel<<<N, threads, smem, mscl::as_stream(stream)>>>(d_xq, d_xf, N, C, t, t2, inv_T,
                                                              d_out, d_gxq, d_gxf, d_part);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

}  // extern "C"
