// K7: trilinear up-sampling of the TPN neck's pyramid levels
// (mmaction/models/necks/sepc.py:126-130: F.interpolate(..., size=[T,H,W], mode="trilinear"), align_corners=False).
//
// Adjacent to the contrastive path: PyTorch's upsample_trilinear3d kernels walk N*C inside
// one thread per output position and run at ~30 GB/s on these shapes (7.5 ms of an 87 ms step); this is a plain
// streaming pass -- one thread per 4 consecutive outputs along W, 128-bit stores, the 8x smaller input served
// from L1/L2 -- and a gather-form backward (no atomics: bit-reproducible).
//
// Source coordinate (ATen area_pixel_compute_source_index, align_corners=False):
//   src = max(0, (dst + 0.5) * in/out - 0.5);  i0 = floor(src);  i1 = min(i0 + 1, in - 1);  l1 = src - i0;  l0 = 1 - l1
#include "common.cuh"

namespace mscl {

struct AxisTap {
  int i0, i1;
  float l0, l1;
};

__device__ __forceinline__ AxisTap axis_tap(int dst, float scale, int in_size) {
  float src = ((float)dst + 0.5f) * scale - 0.5f;
  src = src < 0.f ? 0.f : src;
  AxisTap a;
  a.i0 = (int)src;
  if (a.i0 > in_size - 1) a.i0 = in_size - 1;
  a.i1 = a.i0 + (a.i0 < in_size - 1 ? 1 : 0);
  a.l1 = src - (float)a.i0;
  a.l0 = 1.f - a.l1;
  return a;
}

// x [NC, Ti, Hi, Wi] -> y [NC, To, Ho, Wo]; one thread per 4 outputs along W (Wo % 4 == 0) or per output
template <int VEC>
__global__ void __launch_bounds__(256)
upsample_trilinear_fwd_kernel(const float *__restrict__ x, float *__restrict__ y, int64_t total_v, int Ti, int Hi,
                              int Wi, int To, int Ho, int Wo, float st, float sh, float sw) {
  const int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (v >= total_v) return;
  const int wv = Wo / VEC;
  const int ow0 = (int)(v % wv) * VEC;
  int64_t r = v / wv;
  const int oh = (int)(r % Ho);
  r /= Ho;
  const int ot = (int)(r % To);
  const int64_t nc = r / To;
  const AxisTap at = axis_tap(ot, st, Ti), ah = axis_tap(oh, sh, Hi);
  const float *p = x + nc * (int64_t)Ti * Hi * Wi;
  const float *r00 = p + ((int64_t)at.i0 * Hi + ah.i0) * Wi, *r01 = p + ((int64_t)at.i0 * Hi + ah.i1) * Wi;
  const float *r10 = p + ((int64_t)at.i1 * Hi + ah.i0) * Wi, *r11 = p + ((int64_t)at.i1 * Hi + ah.i1) * Wi;
  float o[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    const AxisTap aw = axis_tap(ow0 + e, sw, Wi);
    // same association as ATen's upsample_trilinear3d_out_frame
    o[e] = at.l0 * (ah.l0 * (aw.l0 * __ldg(r00 + aw.i0) + aw.l1 * __ldg(r00 + aw.i1)) +
                    ah.l1 * (aw.l0 * __ldg(r01 + aw.i0) + aw.l1 * __ldg(r01 + aw.i1))) +
           at.l1 * (ah.l0 * (aw.l0 * __ldg(r10 + aw.i0) + aw.l1 * __ldg(r10 + aw.i1)) +
                    ah.l1 * (aw.l0 * __ldg(r11 + aw.i0) + aw.l1 * __ldg(r11 + aw.i1)));
  }
  float *dst = y + ((nc * To + ot) * (int64_t)Ho + oh) * Wo + ow0;
  if (VEC == 4)
    stg_stream(reinterpret_cast<float4 *>(dst), make_float4(o[0], o[1], o[2], o[3]));
  else
    dst[0] = o[0];
}

// Range of outputs whose taps can touch input index i: src(o) in (i - 1, i + 1)  (plus the clamped ends)
__device__ __forceinline__ void out_range(int i, float inv_scale, int out_size, int &lo, int &hi) {
  lo = (int)floorf(((float)i - 0.5f) * inv_scale - 0.5f) - 1;
  hi = (int)ceilf(((float)i + 1.5f) * inv_scale - 0.5f) + 1;
  lo = lo < 0 ? 0 : lo;
  hi = hi > out_size - 1 ? out_size - 1 : hi;
}
__device__ __forceinline__ float tap_weight(const AxisTap &a, int i) {
  return (a.i0 == i ? a.l0 : 0.f) + (a.i1 == i ? a.l1 : 0.f);
}

// gx [NC, Ti, Hi, Wi] = sum over the outputs that read it; one thread per input element, fixed summation order
__global__ void __launch_bounds__(256)
upsample_trilinear_bwd_kernel(const float *__restrict__ gy, float *__restrict__ gx, int64_t total, int Ti, int Hi,
                              int Wi, int To, int Ho, int Wo, float st, float sh, float sw) {
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= total) return;
  const int iw = (int)(e % Wi);
  int64_t r = e / Wi;
  const int ih = (int)(r % Hi);
  r /= Hi;
  const int it = (int)(r % Ti);
  const int64_t nc = r / Ti;
  int t0, t1, h0, h1, w0, w1;
  out_range(it, 1.f / st, To, t0, t1);
  out_range(ih, 1.f / sh, Ho, h0, h1);
  out_range(iw, 1.f / sw, Wo, w0, w1);
  // the W-axis weights do not depend on (ot, oh): evaluate them once (the candidate range is short: <= 2/scale + 4)
  constexpr int kMaxW = 12;
  float ww[kMaxW];
  const int nw = w1 - w0 + 1;
  const bool cached = nw <= kMaxW;
  if (cached) {
#pragma unroll
    for (int k = 0; k < kMaxW; ++k) ww[k] = (k < nw) ? tap_weight(axis_tap(w0 + k, sw, Wi), iw) : 0.f;
  }
  const float *g = gy + nc * (int64_t)To * Ho * Wo;
  float acc = 0.f;
  for (int ot = t0; ot <= t1; ++ot) {
    const float wt = tap_weight(axis_tap(ot, st, Ti), it);
    if (wt == 0.f) continue;
    for (int oh = h0; oh <= h1; ++oh) {
      const float wh = tap_weight(axis_tap(oh, sh, Hi), ih);
      if (wh == 0.f) continue;
      const float *row = g + ((int64_t)ot * Ho + oh) * Wo + w0;
      float racc = 0.f;
      if (cached) {
#pragma unroll
        for (int k = 0; k < kMaxW; ++k)
          if (k < nw) racc = fmaf(ww[k], __ldg(row + k), racc);
      } else {
        for (int k = 0; k < nw; ++k) racc = fmaf(tap_weight(axis_tap(w0 + k, sw, Wi), iw), __ldg(row + k), racc);
      }
      acc = fmaf(wt * wh, racc, acc);
    }
  }
  gx[e] = acc;
}

// ---- channels-last (NDHWC) forms: the encoders run in torch.channels_last_3d, so the pyramid levels arrive as
// [N, T, H, W, C] with C contiguous.  One thread per (output voxel, 4 channels): the 32 threads of a 128-channel voxel
// read eight 512-byte input rows (L1/L2: the input is 8x smaller than the output) and write one; no layout copy on
// either side, and the sum that follows in the neck stays a same-layout vectorised add.
// One CTA per output row (n, ot, oh): the T/H taps are CTA-uniform, all index arithmetic is 32-bit (the first version
// spent its time in 64-bit div/mod: 41 us for a 58 MB pass), the Wo * C/4 items of the row are strided over the threads.
__global__ void __launch_bounds__(256)
upsample_trilinear_ndhwc_fwd_kernel(const float4 *__restrict__ x, float4 *__restrict__ y, int C4, int Ti, int Hi, int Wi,
                                    int To, int Ho, int Wo, float st, float sh, float sw) {
  const unsigned row = blockIdx.x;                 // (n * To + ot) * Ho + oh
  const int oh = (int)(row % (unsigned)Ho);
  const unsigned r = row / (unsigned)Ho;
  const int ot = (int)(r % (unsigned)To);
  const int64_t n = r / (unsigned)To;
  const AxisTap at = axis_tap(ot, st, Ti), ah = axis_tap(oh, sh, Hi);
  const float4 *p = x + n * (int64_t)Ti * Hi * Wi * C4;
  const float4 *r00 = p + (at.i0 * Hi + ah.i0) * Wi * C4, *r01 = p + (at.i0 * Hi + ah.i1) * Wi * C4;
  const float4 *r10 = p + (at.i1 * Hi + ah.i0) * Wi * C4, *r11 = p + (at.i1 * Hi + ah.i1) * Wi * C4;
  float4 *dst = y + (int64_t)row * Wo * C4;
  const int items = Wo * C4;
  for (int i = threadIdx.x; i < items; i += 256) {
    const int ow = i / C4, cg = i - ow * C4;
    const AxisTap aw = axis_tap(ow, sw, Wi);
    const int o0 = aw.i0 * C4 + cg, o1 = aw.i1 * C4 + cg;
    const float4 a000 = __ldg(r00 + o0), a001 = __ldg(r00 + o1), a010 = __ldg(r01 + o0), a011 = __ldg(r01 + o1),
                 a100 = __ldg(r10 + o0), a101 = __ldg(r10 + o1), a110 = __ldg(r11 + o0), a111 = __ldg(r11 + o1);
    // same association as ATen's upsample_trilinear3d_out_frame
#define MSCL_TRI(f)                                                                                       \
  (at.l0 * (ah.l0 * (aw.l0 * a000.f + aw.l1 * a001.f) + ah.l1 * (aw.l0 * a010.f + aw.l1 * a011.f)) +    \
   at.l1 * (ah.l0 * (aw.l0 * a100.f + aw.l1 * a101.f) + ah.l1 * (aw.l0 * a110.f + aw.l1 * a111.f)))
    stg_stream(dst + i, make_float4(MSCL_TRI(x), MSCL_TRI(y), MSCL_TRI(z), MSCL_TRI(w)));
#undef MSCL_TRI
  }
}

// gx [N, Ti, Hi, Wi, C]: one CTA per input row (n, it, ih); gather over the (at most ~4 per axis) outputs that read
// the voxel, fixed order, no atomics.  The (ot, oh) candidates and their weights are CTA-uniform.
__global__ void __launch_bounds__(256)
upsample_trilinear_ndhwc_bwd_kernel(const float4 *__restrict__ gy, float4 *__restrict__ gx, int C4, int Ti, int Hi,
                                    int Wi, int To, int Ho, int Wo, float st, float sh, float sw) {
  const unsigned row = blockIdx.x;                 // (n * Ti + it) * Hi + ih
  const int ih = (int)(row % (unsigned)Hi);
  const unsigned r = row / (unsigned)Hi;
  const int it = (int)(r % (unsigned)Ti);
  const int64_t n = r / (unsigned)Ti;
  int t0, t1, h0, h1;
  out_range(it, 1.f / st, To, t0, t1);
  out_range(ih, 1.f / sh, Ho, h0, h1);
  const float4 *g = gy + n * (int64_t)To * Ho * Wo * C4;
  float4 *dst = gx + (int64_t)row * Wi * C4;
  const int items = Wi * C4;
  const float inv_sw = 1.f / sw;
  for (int i = threadIdx.x; i < items; i += 256) {
    const int iw = i / C4, cg = i - iw * C4;
    int w0, w1;
    out_range(iw, inv_sw, Wo, w0, w1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ot = t0; ot <= t1; ++ot) {
      const float wt = tap_weight(axis_tap(ot, st, Ti), it);
      if (wt == 0.f) continue;
      for (int oh = h0; oh <= h1; ++oh) {
        const float wth = wt * tap_weight(axis_tap(oh, sh, Hi), ih);
        if (wth == 0.f) continue;
        const float4 *grow = g + (ot * Ho + oh) * Wo * C4 + cg;
        for (int ow = w0; ow <= w1; ++ow) {
          const float wgt = wth * tap_weight(axis_tap(ow, sw, Wi), iw);
          if (wgt == 0.f) continue;
          const float4 q = __ldg(grow + ow * C4);      // neighbouring voxels of the CTA share taps: keep L1
          acc.x = fmaf(wgt, q.x, acc.x);
          acc.y = fmaf(wgt, q.y, acc.y);
          acc.z = fmaf(wgt, q.z, acc.z);
          acc.w = fmaf(wgt, q.w, acc.w);
        }
      }
    }
    dst[i] = acc;
  }
}

// Separable backward for channels-last tensors: the trilinear weights factorise, so the gradient of the up-sampling is
// three 1-D linear-interpolation backward passes (W, then H, then T), each a gather over the <= ~4 outputs that read an
// input index along that axis.  src [outer][out_size][inner4] -> dst [outer][in_size][inner4] in float4 units; every
// pass reads its input once, fully coalesced (the direct 3-D gather re-reads each output row for ~4 input rows and
// is L2-bound: 72 us for a 58 MB pass at the r18 sizes).
__global__ void __launch_bounds__(256)
linear_axis_bwd_kernel(const float4 *__restrict__ src, float4 *__restrict__ dst, unsigned total, int in_size, int out_size,
                       int inner4, float scale, float inv_scale) {
  const unsigned e = blockIdx.x * 256u + threadIdx.x;
  if (e >= total) return;
  const unsigned r = e / (unsigned)inner4;
  const int inner = (int)(e - r * (unsigned)inner4);
  const unsigned outer = r / (unsigned)in_size;
  const int i = (int)(r - outer * (unsigned)in_size);
  int o0, o1;
  out_range(i, inv_scale, out_size, o0, o1);
  const float4 *p = src + ((int64_t)outer * out_size) * inner4 + inner;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int o = o0; o <= o1; ++o) {
    const float w = tap_weight(axis_tap(o, scale, in_size), i);
    if (w == 0.f) continue;
    const float4 q = __ldg(p + (int64_t)o * inner4);
    acc.x = fmaf(w, q.x, acc.x);
    acc.y = fmaf(w, q.y, acc.y);
    acc.z = fmaf(w, q.z, acc.z);
    acc.w = fmaf(w, q.w, acc.w);
  }
  dst[e] = acc;
}

}  // namespace mscl

extern "C" {

static int resample_check(const void *a, const void *b, int64_t NC, int Ti, int Hi, int Wi, int To, int Ho, int Wo) {
  MSCL_CHECK_ARG(a && b, "null pointer");
  MSCL_CHECK_ARG(NC > 0 && Ti > 0 && Hi > 0 && Wi > 0 && To > 0 && Ho > 0 && Wo > 0, "bad shape");
  MSCL_CHECK_ARG((((uintptr_t)a | (uintptr_t)b) & 15) == 0, "tensors must be 16-byte aligned");
  return MSCL_OK;
}

int mscl_upsample_trilinear_fwd(const float *d_x, float *d_y, int64_t NC, int32_t Ti, int32_t Hi, int32_t Wi,
                                int32_t To, int32_t Ho, int32_t Wo, mscl_stream_t stream) {
  int rc = resample_check(d_x, d_y, NC, Ti, Hi, Wi, To, Ho, Wo);
  if (rc) return rc;
  const float st = (float)Ti / (float)To, sh = (float)Hi / (float)Ho, sw = (float)Wi / (float)Wo;
  const int vec = (Wo % 4 == 0) ? 4 : 1;
  const int64_t total_v = NC * To * Ho * (Wo / vec);
  const int64_t blocks = (total_v + 255) / 256;
  MSCL_CHECK_ARG(blocks < (1ll << 31), "too many elements");
  cudaStream_t s = mscl::as_stream(stream);
  if (vec == 4)
    mscl::upsample_trilinear_fwd_kernel<4><<<(unsigned)blocks, 256, 0, s>>>(d_x, d_y, total_v, Ti, Hi, Wi, To, Ho, Wo, st, sh, sw);
  else
    mscl::upsample_trilinear_fwd_kernel<1><<<(unsigned)blocks, 256, 0, s>>>(d_x, d_y, total_v, Ti, Hi, Wi, To, Ho, Wo, st, sh, sw);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

int mscl_upsample_trilinear_bwd(const float *d_gy, float *d_gx, int64_t NC, int32_t Ti, int32_t Hi, int32_t Wi,
                                int32_t To, int32_t Ho, int32_t Wo, mscl_stream_t stream) {
  int rc = resample_check(d_gy, d_gx, NC, Ti, Hi, Wi, To, Ho, Wo);
  if (rc) return rc;
  const float st = (float)Ti / (float)To, sh = (float)Hi / (float)Ho, sw = (float)Wi / (float)Wo;
  const int64_t total = NC * Ti * Hi * Wi;
  const int64_t blocks = (total + 255) / 256;
  MSCL_CHECK_ARG(blocks < (1ll << 31), "too many elements");
  mscl::upsample_trilinear_bwd_kernel<<<(unsigned)blocks, 256, 0, mscl::as_stream(stream)>>>(d_gy, d_gx, total, Ti, Hi, Wi, To,
                                                                                            Ho, Wo, st, sh, sw);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

int mscl_upsample_trilinear_ndhwc_fwd(const float *d_x, float *d_y, int64_t N, int32_t C, int32_t Ti, int32_t Hi, int32_t Wi,
                                      int32_t To, int32_t Ho, int32_t Wo, mscl_stream_t stream) {
  int rc = resample_check(d_x, d_y, N, Ti, Hi, Wi, To, Ho, Wo);
  if (rc) return rc;
  MSCL_CHECK_ARG(C > 0 && C % 4 == 0, "channels-last form needs C %% 4 == 0 (C=%d)", C);
  const float st = (float)Ti / (float)To, sh = (float)Hi / (float)Ho, sw = (float)Wi / (float)Wo;
  const int64_t rows = N * To * Ho;
  MSCL_CHECK_ARG(rows < (1ll << 31) && (int64_t)Ti * Hi * Wi * (C / 4) < (1ll << 31) && (int64_t)To * Ho * Wo * (C / 4) < (1ll << 31),
                 "sample too large for 32-bit indexing");
  mscl::upsample_trilinear_ndhwc_fwd_kernel<<<(unsigned)rows, 256, 0, mscl::as_stream(stream)>>>(
      reinterpret_cast<const float4 *>(d_x), reinterpret_cast<float4 *>(d_y), C / 4, Ti, Hi, Wi, To, Ho, Wo, st, sh, sw);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

int mscl_upsample_trilinear_ndhwc_bwd(const float *d_gy, float *d_gx, int64_t N, int32_t C, int32_t Ti, int32_t Hi, int32_t Wi,
                                      int32_t To, int32_t Ho, int32_t Wo, mscl_stream_t stream) {
  int rc = resample_check(d_gy, d_gx, N, Ti, Hi, Wi, To, Ho, Wo);
  if (rc) return rc;
  MSCL_CHECK_ARG(C > 0 && C % 4 == 0, "channels-last form needs C %% 4 == 0 (C=%d)", C);
  const float st = (float)Ti / (float)To, sh = (float)Hi / (float)Ho, sw = (float)Wi / (float)Wo;
  const int64_t rows = N * Ti * Hi;
  MSCL_CHECK_ARG(rows < (1ll << 31) && (int64_t)Ti * Hi * Wi * (C / 4) < (1ll << 31) && (int64_t)To * Ho * Wo * (C / 4) < (1ll << 31),
                 "sample too large for 32-bit indexing");
  mscl::upsample_trilinear_ndhwc_bwd_kernel<<<(unsigned)rows, 256, 0, mscl::as_stream(stream)>>>(
      reinterpret_cast<const float4 *>(d_gy), reinterpret_cast<float4 *>(d_gx), C / 4, Ti, Hi, Wi, To, Ho, Wo, st, sh, sw);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

int mscl_linear_axis_bwd(const float *d_src, float *d_dst, int64_t outer, int32_t in_size, int32_t out_size, int64_t inner4,
                         mscl_stream_t stream) {
  MSCL_CHECK_ARG(d_src && d_dst, "null pointer");
  MSCL_CHECK_ARG(outer > 0 && in_size > 0 && out_size > 0 && inner4 > 0, "bad shape");
  MSCL_CHECK_ARG((((uintptr_t)d_src | (uintptr_t)d_dst) & 15) == 0, "tensors must be 16-byte aligned");
  const int64_t total = outer * in_size * inner4;
  MSCL_CHECK_ARG(total < (1ll << 32) && inner4 < (1ll << 31) && outer * out_size * inner4 < (1ll << 40), "tensor too large");
  const float scale = (float)in_size / (float)out_size;
  mscl::linear_axis_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, mscl::as_stream(stream)>>>(
      reinterpret_cast<const float4 *>(d_src), reinterpret_cast<float4 *>(d_dst), (unsigned)total, in_size, out_size, (int)inner4,
      scale, 1.f / scale);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

}  // extern "C"
