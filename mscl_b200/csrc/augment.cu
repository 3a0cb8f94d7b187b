// K8 / K9: the GPU augmentation front-end that feeds the encoders every step.
//
//  K8 flow_visualize : (u,v) flow -> Middlebury colour-wheel image, the reference's FlowVisualizer
//                      (mmaction/models/common/ssl_aug.py:87-136, wheel tools/RAFT/core/utils/flow_viz.py:20-67),
//                      fused with the per-sample horizontal flip of SyncMoCoAugmentV5.forward_flip
//                      (common/ssl_aug_v2.py:109-117: only the image is mirrored, u keeps its sign) and the optional
//                      normalisation.  One thread per 4 pixels: 8 B read, 12 B written per pixel.
//  K9 color_pipeline : the RGB branch of SyncMoCoAugmentV5 (ssl_aug_v2.py:31-48,66-68): flip, ColorJitter
//                      (brightness, contrast, saturation, hue) p=.8, RandomGrayscale p=.2, Gaussian blur p=.5
//                      (separable, reflect border), Normalize -- decisions and parameters per clip, drawn by the
//                      caller.  clip_gray_sum (the contrast step needs the clip's mean luminance) + one CTA per
//                      frame that keeps the frame in shared memory through both blur passes: 12 B read (x2 for the
//                      mean) and 12 B written per pixel instead of ~25 element-wise passes and two convolutions.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace mscl {

// ------------------------------------------------------------------------------------------- K8
__constant__ double c_wheel[55 * 3];        // colour wheel / 255 (float64, as the reference's tmp[k] / 255.0)

struct WheelTap {
  int k0, k1;
  float f, omf, rad;
};

__device__ __forceinline__ WheelTap wheel_tap(float u, float v) {
  WheelTap w;
  w.rad = __fsqrt_rn(__fadd_rn(__fmul_rn(u, u), __fmul_rn(v, v)));
  const float a = __fdiv_rn(atan2f(-v, -u), 3.14159265358979323846f);
  const float fk = __fmul_rn(__fmul_rn(__fadd_rn(a, 1.f), 0.5f), 54.f);      // x / 2 == x * 0.5 exactly
  const float fl = floorf(fk);
  w.k0 = (int)fl;
  w.k1 = w.k0 + 1;
  if (w.k1 == 55) w.k1 = 0;
  w.f = __fsub_rn(fk, fl);
  w.omf = __fsub_rn(1.f, w.f);      // (1 - f) is a float32 op in the reference; the products below are float64
  return w;
}

// `wheel`: the shared-memory copy (the lookup index differs per lane: constant memory would serialise it)
// `div255[k]` = float32(k / 255) for k = 0..255, filled once per CTA with the correctly rounded division: the kernel is
// instruction-bound (~300 per pixel), and three IEEE divisions per pixel were a third of that.
__device__ __forceinline__ float wheel_color(const WheelTap &w, int ch, const double *wheel, const float *div255) {
  double col = __dadd_rn(__dmul_rn((double)w.omf, wheel[w.k0 * 3 + ch]), __dmul_rn((double)w.f, wheel[w.k1 * 3 + ch]));
  if (w.rad <= 1.f)
    col = __dsub_rn(1.0, __dmul_rn((double)w.rad, __dsub_rn(1.0, col)));
  else
    col = __dmul_rn(col, 0.75);
  const unsigned char q = (unsigned char)floor(__dmul_rn(255.0, col));     // uint8 round trip (ssl_aug.py:121)
  return div255[q];
}

// flow planar [N, 2, T, H, W] -> out [N, 3, T, H, W].  flip: uint8 [N] or null.  norm: float [6] = mean[3], std[3] or null.
__global__ void __launch_bounds__(256)
flow_visualize_kernel(const float *__restrict__ flow, const uint8_t *__restrict__ flip, const float *__restrict__ norm,
                      float *__restrict__ out, int T, int H, int W, int64_t total4) {
  __shared__ double s_wheel[55 * 3];
  __shared__ float s_div255[256];
  if (threadIdx.x < 55 * 3) s_wheel[threadIdx.x] = c_wheel[threadIdx.x];
  s_div255[threadIdx.x] = __fdiv_rn((float)threadIdx.x, 255.f);             // blockDim.x == 256
  __syncthreads();
  const unsigned g = blockIdx.x * 256u + threadIdx.x;       // total4 < 2^31 (checked by the launcher): 32-bit div / mod
  if (g >= total4) return;
  const unsigned w4 = (unsigned)W >> 2;
  const unsigned r = g / w4;                                 // (n, t, h) row index
  const int wq = (int)(g - r * w4);
  const int64_t HW = (int64_t)H * W, THW = (int64_t)T * HW;
  const unsigned th = (unsigned)T * (unsigned)H;
  const int n = (int)(r / th);
  const int64_t row_in_clip = (int64_t)(r - (unsigned)n * th);
  const bool fl = flip != nullptr && flip[n] != 0;
  const int w_src = fl ? (W - 4 - 4 * wq) : 4 * wq;
  const float *pu = flow + ((int64_t)n * 2 + 0) * THW + row_in_clip * W + w_src;
  const float *pv = flow + ((int64_t)n * 2 + 1) * THW + row_in_clip * W + w_src;
  float4 u4 = ldg_stream(reinterpret_cast<const float4 *>(pu));
  float4 v4 = ldg_stream(reinterpret_cast<const float4 *>(pv));
  if (fl) {
    u4 = make_float4(u4.w, u4.z, u4.y, u4.x);
    v4 = make_float4(v4.w, v4.z, v4.y, v4.x);
  }
  const float us[4] = {u4.x, u4.y, u4.z, u4.w}, vs[4] = {v4.x, v4.y, v4.z, v4.w};
  float o[3][4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const WheelTap tap = wheel_tap(us[e], vs[e]);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float c = wheel_color(tap, ch, s_wheel, s_div255);
      if (norm != nullptr) c = __fdiv_rn(__fsub_rn(c, norm[ch]), norm[3 + ch]);
      o[ch][e] = c;
    }
  }
  float *po = out + (int64_t)n * 3 * THW + row_in_clip * W + 4 * wq;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch)
    stg_stream(reinterpret_cast<float4 *>(po + ch * THW), make_float4(o[ch][0], o[ch][1], o[ch][2], o[ch][3]));
}

// ------------------------------------------------------------------------------------------- K9
__device__ __forceinline__ float gray_of(float r, float g, float b) {
  return 0.299f * r + 0.587f * g + 0.114f * b;
}

// partial[u][chunk] = sum over this chunk of unit u's pixels of gray(x); grid (chunks, units).  A unit is a clip
// (frames_per_unit = T: `count` = T*H*W pixels) or one frame of it (frames_per_unit = 1: `count` = H*W pixels).
__global__ void __launch_bounds__(256)
clip_gray_sum_kernel(const float *__restrict__ x, float *__restrict__ partial, int64_t THW, int64_t count, int units_per_clip) {
  const int u = blockIdx.y;
  const int n = u / units_per_clip, f = u - n * units_per_clip;
  const float *pr = x + (int64_t)n * 3 * THW + (int64_t)f * count, *pg = pr + THW, *pb = pg + THW;
  const int64_t n4 = count >> 2;
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const float4 r = ldg_stream(reinterpret_cast<const float4 *>(pr) + i);
    const float4 g = ldg_stream(reinterpret_cast<const float4 *>(pg) + i);
    const float4 b = ldg_stream(reinterpret_cast<const float4 *>(pb) + i);
    acc += (gray_of(r.x, g.x, b.x) + gray_of(r.y, g.y, b.y)) + (gray_of(r.z, g.z, b.z) + gray_of(r.w, g.w, b.w));
  }
  acc = warp_sum(acc);
  __shared__ float s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s[w];
    partial[u * gridDim.x + blockIdx.x] = t;
  }
}

// Parameters, float [N][kColorParams] (per clip) or [N*T][kColorParams] (per frame: the reference's 'batch' sync level draws
// the jitter factors per image, ssl_aug.py:33-60; the contrast step then uses the frame's own mean luminance):
//  0 flip  1 jitter?  2 brightness  3 contrast  4 saturation  5..13 hue matrix (row major)  14 gray?  15 blur?
constexpr int kColorParams = 16;
constexpr int kMaxTaps = 31;

// One CTA per frame (n, t).  smem: a[H*W] | b[H*W] floats.
__global__ void __launch_bounds__(512)
color_pipeline_kernel(const float *__restrict__ x, const float *__restrict__ params, const float *__restrict__ gray_partial,
                      int n_chunks, const float *__restrict__ taps, int n_taps, const float *__restrict__ norm,
                      float *__restrict__ out, int T, int H, int W, int per_frame) {
  extern __shared__ float sm[];
  const int HW = H * W;
  float *a = sm, *b = sm + HW;
  __shared__ float s_taps[kMaxTaps];
  const int frame = blockIdx.x;
  const int n = frame / T, t = frame - n * T;
  const int unit = per_frame ? frame : n;
  const float *p = params + unit * kColorParams;
  const bool flip = p[0] != 0.f, jit = p[1] != 0.f, to_gray = p[14] != 0.f, blur = p[15] != 0.f;
  const float br = p[2], ct = p[3], sat = p[4];
  float hm[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) hm[i] = p[5 + i];
  const int64_t THW = (int64_t)T * HW;
  // clip mean of gray(x * brightness): the partial sums in a fixed order
  float gsum = 0.f;
  for (int c = 0; c < n_chunks; ++c) gsum += __ldg(gray_partial + unit * n_chunks + c);
  const float m = br * (gsum / (per_frame ? (float)HW : (float)THW));     // mean luminance of the clip / of the frame
  if (threadIdx.x < n_taps) s_taps[threadIdx.x] = taps[threadIdx.x];
  const float *pr = x + (int64_t)n * 3 * THW + (int64_t)t * HW, *pg = pr + THW, *pb = pg + THW;
  float *po = out + (int64_t)n * 3 * THW + (int64_t)t * HW;
  const int half = n_taps >> 1;
  // point-wise colour ops of one pixel (flip applied on the way in)
  auto shade = [&](int i, float &r, float &g, float &bl) {
    const int h = i / W, w = i - h * W;
    const int src = flip ? (h * W + (W - 1 - w)) : i;
    r = __ldg(pr + src), g = __ldg(pg + src), bl = __ldg(pb + src);
    if (jit) {
      float y0 = r * br, y1 = g * br, y2 = bl * br;                       // brightness
      y0 = (y0 - m) * ct + m, y1 = (y1 - m) * ct + m, y2 = (y2 - m) * ct + m;   // contrast about the clip mean
      const float gy = gray_of(y0, y1, y2);
      y0 = (y0 - gy) * sat + gy, y1 = (y1 - gy) * sat + gy, y2 = (y2 - gy) * sat + gy;   // saturation
      const float z0 = hm[0] * y0 + hm[1] * y1 + hm[2] * y2;             // hue rotation in YIQ space
      const float z1 = hm[3] * y0 + hm[4] * y1 + hm[5] * y2;
      const float z2 = hm[6] * y0 + hm[7] * y1 + hm[8] * y2;
      r = fminf(fmaxf(z0, 0.f), 1.f), g = fminf(fmaxf(z1, 0.f), 1.f), bl = fminf(fmaxf(z2, 0.f), 1.f);
    }
    if (to_gray) r = g = bl = gray_of(r, g, bl);
  };
  if (!blur) {      // one pass, three planes out
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      float r, g, bl;
      shade(i, r, g, bl);
      po[i] = (r - norm[0]) / norm[3];
      po[THW + i] = (g - norm[1]) / norm[4];
      po[2 * THW + i] = (bl - norm[2]) / norm[5];
    }
    return;
  }
  __syncthreads();   // s_taps
  for (int ch = 0; ch < 3; ++ch) {
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      float r, g, bl;
      shade(i, r, g, bl);
      a[i] = ch == 0 ? r : (ch == 1 ? g : bl);
    }
    __syncthreads();
    // ---- horizontal pass a -> b, reflect border (F.pad mode="reflect": index -k -> k, W-1+k -> W-1-k)
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      const int h = i / W, w = i - h * W;
      float acc = 0.f;
      for (int k = 0; k < n_taps; ++k) {
        int ww = w + k - half;
        ww = ww < 0 ? -ww : (ww >= W ? 2 * W - 2 - ww : ww);
        acc = fmaf(s_taps[k], a[h * W + ww], acc);
      }
      b[i] = acc;
    }
    __syncthreads();
    // ---- vertical pass b -> out, normalised
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      const int h = i / W, w = i - h * W;
      float acc = 0.f;
      for (int k = 0; k < n_taps; ++k) {
        int hh = h + k - half;
        hh = hh < 0 ? -hh : (hh >= H ? 2 * H - 2 - hh : hh);
        acc = fmaf(s_taps[k], b[hh * W + w], acc);
      }
      po[ch * THW + i] = (acc - norm[ch]) / norm[3 + ch];
    }
    __syncthreads();
  }
}

// Fast path for frames up to 128 pixels wide (the config's 112x112 crops).  A CTA owns one BAND of rows of one frame
// (plus TAPS/2 halo rows each side): the band's three planes are staged in shared memory by bulk async copies, shaded
// in place once per pixel, blurred horizontally IN PLACE (a warp owns a row: every lane loads its 4 + TAPS - 1 inputs,
// __syncwarp, stores 4 outputs), and the vertical pass streams 8 rows per thread from shared memory straight to
// global.  Bands keep the footprint at ~65 KB, so three CTAs share an SM and one band's loads overlap another's
// arithmetic and stores.  ~180 instructions per pixel instead of ~870 in the generic kernel.
// The mirror of a flipped clip is applied when storing (the blur is symmetric, so flipping commutes with it).
// CLUSTER_MEAN (per-frame parameters): the bands of a frame form a thread-block cluster; each CTA sums the luminance of its own
// rows out of the staged (raw) planes and reads its peers' sums through distributed shared memory, so the frame's mean
// luminance -- what the contrast step is taken about -- needs no pass over the clip before this kernel (clip_gray_sum read
// the 38 MB input a second time: 11 of K9's 47 us per view).
template <int TAPS, bool CLUSTER_MEAN>
__global__ void __launch_bounds__(512)
color_pipeline_fast_kernel(const float *__restrict__ x, const float *__restrict__ params, const float *__restrict__ gray_partial,
                           int n_chunks, const float *__restrict__ taps, const float *__restrict__ norm,
                           float *__restrict__ out, int T, int H, int W, int band_rows, int per_frame) {
  extern __shared__ float sm[];
  constexpr int HALF = TAPS / 2;
  const int HW = H * W;
  const int frame = blockIdx.x;
  const int n = frame / T, t = frame - n * T;
  const int unit = per_frame ? frame : n;
  const int r0 = blockIdx.y * band_rows, r1 = min(r0 + band_rows, H);        // output rows of this band
  if (!CLUSTER_MEAN && r0 >= H) return;      // (the cluster form is only launched when every band has rows)
  const float *p = params + unit * kColorParams;
  const bool flip = p[0] != 0.f, jit = p[1] != 0.f, to_gray = p[14] != 0.f, blur = p[15] != 0.f;
  const int s0 = blur ? max(r0 - HALF, 0) : r0, s1 = blur ? min(r1 + HALF, H) : r1;   // staged rows
  const int SR = s1 - s0, SP = SR * W;                                         // rows / floats per staged plane
  const float br = p[2], ct = p[3], sat = p[4];
  float hm[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) hm[i] = p[5 + i];
  float tp[TAPS];
#pragma unroll
  for (int k = 0; k < TAPS; ++k) tp[k] = __ldg(taps + k);
  const int64_t THW = (int64_t)T * HW;
  float gsum = 0.f;
  if (!CLUSTER_MEAN)
    for (int c = 0; c < n_chunks; ++c) gsum += __ldg(gray_partial + unit * n_chunks + c);
  float m = br * (gsum / (per_frame ? (float)HW : (float)THW));
  const float *pr = x + (int64_t)n * 3 * THW + (int64_t)t * HW;
  float *po = out + (int64_t)n * 3 * THW + (int64_t)t * HW;
  const float n0 = norm[0], n1 = norm[1], n2 = norm[2];
  const int nthr = blockDim.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = nthr >> 5;
  // ---- phase 0: staged rows of the three planes -> shared memory (bulk async copies, one mbarrier)
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"((uint32_t)(3 * SP * 4)) : "memory");
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm);
#pragma unroll
    for (int c = 0; c < 3; ++c)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       dst + (uint32_t)(c * SP * 4)), "l"(pr + c * THW + (int64_t)s0 * W), "r"((uint32_t)(SP * 4)), "r"(bar_a)
                   : "memory");
  }
  __syncthreads();
  {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar_a) : "memory");
  }
  if (CLUSTER_MEAN && jit) {      // `jit` is a property of the frame: the whole cluster takes this branch or none of it does
    __shared__ float cl_part[16];
    __shared__ float cl_sum;
    const float4 *q4 = reinterpret_cast<const float4 *>(sm + (r0 - s0) * W);      // this band's own rows, plane 0
    const int own4 = (r1 - r0) * W >> 2, plane4 = SP >> 2;
    float acc = 0.f;
    for (int i = threadIdx.x; i < own4; i += nthr) {
      const float4 r = q4[i], g = q4[plane4 + i], b = q4[2 * plane4 + i];
      acc += (gray_of(r.x, g.x, b.x) + gray_of(r.y, g.y, b.y)) + (gray_of(r.z, g.z, b.z) + gray_of(r.w, g.w, b.w));
    }
    acc = warp_sum(acc);
    if (lane == 0) cl_part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < nwarps; ++i) t += cl_part[i];
      cl_sum = t;
    }
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();
    float total = 0.f;
    for (int rk = 0; rk < (int)cluster.num_blocks(); ++rk) total += *cluster.map_shared_rank(&cl_sum, rk);
    m = br * (total / (float)HW);
    cluster.sync();                 // nobody leaves (or moves on to overwrite anything) while a peer may still read its sum
  }
  // ---- phase 1: shade every staged pixel once, in place, four pixels per thread (W % 4 == 0 -> SP % 4 == 0)
  auto shade1 = [&](float &r, float &g, float &bl) {
    if (jit) {
      float y0 = r * br, y1 = g * br, y2 = bl * br;
      y0 = (y0 - m) * ct + m, y1 = (y1 - m) * ct + m, y2 = (y2 - m) * ct + m;
      const float gy = gray_of(y0, y1, y2);
      y0 = (y0 - gy) * sat + gy, y1 = (y1 - gy) * sat + gy, y2 = (y2 - gy) * sat + gy;
      const float z0 = hm[0] * y0 + hm[1] * y1 + hm[2] * y2;
      const float z1 = hm[3] * y0 + hm[4] * y1 + hm[5] * y2;
      const float z2 = hm[6] * y0 + hm[7] * y1 + hm[8] * y2;
      r = fminf(fmaxf(z0, 0.f), 1.f), g = fminf(fmaxf(z1, 0.f), 1.f), bl = fminf(fmaxf(z2, 0.f), 1.f);
    }
    if (to_gray) r = g = bl = gray_of(r, g, bl);
  };
  const float rs0 = 1.f / norm[3], rs1 = 1.f / norm[4], rs2 = 1.f / norm[5];
  float4 *sm4 = reinterpret_cast<float4 *>(sm);
  const int SP4 = SP >> 2, W4 = W >> 2;
  for (int i = threadIdx.x; i < SP4; i += nthr) {
    float4 r = sm4[i], g = sm4[SP4 + i], bl = sm4[2 * SP4 + i];
    shade1(r.x, g.x, bl.x);
    shade1(r.y, g.y, bl.y);
    shade1(r.z, g.z, bl.z);
    shade1(r.w, g.w, bl.w);
    if (blur) {
      sm4[i] = r, sm4[SP4 + i] = g, sm4[2 * SP4 + i] = bl;
    } else {      // not blurred: normalise and store straight away (mirrored within the row when flipped)
      const int hh = i / W4, wq = i - hh * W4;
      const int64_t o = (int64_t)(s0 + hh) * W + (flip ? (W - 4 - 4 * wq) : 4 * wq);
      if (flip) {
        r = make_float4(r.w, r.z, r.y, r.x), g = make_float4(g.w, g.z, g.y, g.x), bl = make_float4(bl.w, bl.z, bl.y, bl.x);
      }
      stg_stream(reinterpret_cast<float4 *>(po + o),
                 make_float4((r.x - n0) * rs0, (r.y - n0) * rs0, (r.z - n0) * rs0, (r.w - n0) * rs0));
      stg_stream(reinterpret_cast<float4 *>(po + THW + o),
                 make_float4((g.x - n1) * rs1, (g.y - n1) * rs1, (g.z - n1) * rs1, (g.w - n1) * rs1));
      stg_stream(reinterpret_cast<float4 *>(po + 2 * THW + o),
                 make_float4((bl.x - n2) * rs2, (bl.y - n2) * rs2, (bl.z - n2) * rs2, (bl.w - n2) * rs2));
    }
  }
  if (!blur) return;
  __syncthreads();
  // ---- phase 2: horizontal pass, in place, one warp per staged row of one plane (W <= 128: one segment of 4 per lane)
  for (int task = warp; task < 3 * SR; task += nwarps) {
    float *row = sm + task * W;                 // planes are contiguous: row `task` of the 3*SR stacked rows
    const int w0 = lane * 4;
    // the lane's 4 + TAPS - 1 inputs as whole 16-byte chunks around its own (conflict-free 128-bit shared loads; one scalar
    // load per input is a 4-way bank conflict, the lanes being 4 floats apart), chunks clamped to the row and the few
    // inputs beyond its ends (reflect border: lanes at the row's ends only) patched afterwards
    constexpr int CH = (HALF + 3) / 4;          // chunks needed on each side of the lane's own
    constexpr int OFF = 4 * CH - HALF;          // win[j] = buf[j + OFF]
    float buf[4 * (2 * CH + 1)];
    if (w0 < W) {
      const float4 *row4 = reinterpret_cast<const float4 *>(row);
#pragma unroll
      for (int c = 0; c < 2 * CH + 1; ++c) {
        const int q = min(max(lane - CH + c, 0), W4 - 1);
        const float4 v = row4[q];
        buf[4 * c] = v.x, buf[4 * c + 1] = v.y, buf[4 * c + 2] = v.z, buf[4 * c + 3] = v.w;
      }
      if (w0 < HALF) {            // only the first HALF inputs of a lane can lie left of the row ...
#pragma unroll
        for (int j = 0; j < HALF; ++j) {
          const int ww = w0 - HALF + j;
          if (ww < 0) buf[j + OFF] = row[-ww];
        }
      }
      if (w0 + 3 + HALF >= W) {   // ... and only the last HALF right of it
#pragma unroll
        for (int j = 4 + TAPS - 1 - HALF; j < 4 + TAPS - 1; ++j) {
          const int ww = w0 - HALF + j;
          if (ww >= W) buf[j + OFF] = row[2 * W - 2 - ww];
        }
      }
    }
    __syncwarp();
    if (w0 < W) {
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < TAPS; ++k) acc = fmaf(tp[k], buf[e + k + OFF], acc);
        o[e] = acc;
      }
      *reinterpret_cast<float4 *>(row + w0) = make_float4(o[0], o[1], o[2], o[3]);      // W % 4 == 0: w0 + 3 < W
    }
    __syncwarp();
  }
  __syncthreads();
  // ---- phase 3: vertical pass, 8 output rows per thread, shared -> global, normalised (frame-border rows reflect)
  // thread <-> column (W <= 128: the first 128 threads' lane index), the 4 groups of 128 threads take the
  // (plane, row chunk) tasks in turn: no per-task division by W, and a chunk whose window needs neither the frame's
  // reflect border nor the clamp of a ragged end walks its column with one pointer increment per row
  constexpr int RC = 8;
  const int hchunks = (r1 - r0 + RC - 1) / RC;
  const int w = threadIdx.x & 127, part = threadIdx.x >> 7, nparts = nthr >> 7;
  if (w < W) {
    const int wo = flip ? (W - 1 - w) : w;
    for (int task = part; task < 3 * hchunks; task += nparts) {
      const int ch = task / hchunks, hc = task - ch * hchunks;
      const float *plane = sm + ch * SP;
      const int h0 = r0 + hc * RC;
      float win[RC + TAPS - 1];
      if (h0 - HALF >= s0 && h0 + RC - 1 + HALF < s1) {      // interior: rows h0 - HALF .. h0 + RC - 1 + HALF are all staged
        const float *pc = plane + (h0 - HALF - s0) * W + w;
#pragma unroll
        for (int j = 0; j < RC + TAPS - 1; ++j, pc += W) win[j] = *pc;
      } else {
#pragma unroll
        for (int j = 0; j < RC + TAPS - 1; ++j) {
          int hh = h0 - HALF + j;
          hh = hh < 0 ? -hh : (hh >= H ? 2 * H - 2 - hh : hh);
          hh = min(max(hh, s0), s1 - 1);          // only rows of a ragged last chunk (never stored) get clamped
          win[j] = plane[(hh - s0) * W + w];
        }
      }
      const float mean = ch == 0 ? n0 : (ch == 1 ? n1 : n2);
      const float rs = ch == 0 ? rs0 : (ch == 1 ? rs1 : rs2);
      float *dst = po + ch * THW + (int64_t)h0 * W + wo;
#pragma unroll
      for (int e = 0; e < RC; ++e) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < TAPS; ++k) acc = fmaf(tp[k], win[e + k], acc);
        if (h0 + e < r1) dst[e * W] = (acc - mean) * rs;
      }
    }
  }
}

}  // namespace mscl

extern "C" {

int mscl_flow_visualize(const float *d_flow, const uint8_t *d_flip, const float *d_norm, float *d_out, int32_t N,
                        int32_t T, int32_t H, int32_t W, mscl_stream_t stream) {
  MSCL_CHECK_ARG(d_flow && d_out, "null pointer");
  MSCL_CHECK_ARG(N > 0 && T > 0 && H > 0 && W > 0 && W % 4 == 0, "bad shape N=%d T=%d H=%d W=%d (W must be a multiple of 4)", N,
                 T, H, W);
  MSCL_CHECK_ARG((((uintptr_t)d_flow | (uintptr_t)d_out) & 15) == 0, "flow / out must be 16-byte aligned");
  static bool wheel_ready = false;
  if (!wheel_ready) {
    // 55-entry Middlebury wheel (tools/RAFT/core/utils/flow_viz.py:33-67): six hue segments, floor(255*i/len) ramps
    const int seg_len[6] = {15, 6, 4, 11, 13, 6};
    const int full[6] = {0, 1, 1, 2, 2, 0}, ramp[6] = {1, 0, 2, 1, 0, 2}, rising[6] = {1, 0, 1, 0, 1, 0};
    double wheel[55 * 3];
    int k = 0;
    for (int s = 0; s < 6; ++s)
      for (int i = 0; i < seg_len[s]; ++i, ++k) {
        double rgb[3] = {0.0, 0.0, 0.0};
        const double r = (double)((255 * i) / seg_len[s]);
        rgb[full[s]] = 255.0;
        rgb[ramp[s]] = rising[s] ? r : 255.0 - r;
        for (int c = 0; c < 3; ++c) wheel[k * 3 + c] = rgb[c] / 255.0;
      }
    MSCL_CUDA(cudaMemcpyToSymbol(mscl::c_wheel, wheel, sizeof(wheel)));
    wheel_ready = true;
  }
  const int64_t total4 = (int64_t)N * T * H * (W / 4);
  const int64_t blocks = (total4 + 255) / 256;
  MSCL_CHECK_ARG(total4 < (1ll << 31), "too many pixels");
  mscl::flow_visualize_kernel<<<(unsigned)blocks, 256, 0, mscl::as_stream(stream)>>>(d_flow, d_flip, d_norm, d_out, T, H, W,
                                                                                    total4);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

int mscl_color_pipeline(const float *d_x, const float *d_params, const float *d_taps, int32_t n_taps, const float *d_norm,
                        float *d_gray_partial, int32_t n_chunks, float *d_out, int32_t N, int32_t T, int32_t H, int32_t W,
                        int32_t per_frame, mscl_stream_t stream) {
  MSCL_CHECK_ARG(d_x && d_params && d_taps && d_norm && d_gray_partial && d_out, "null pointer");
  MSCL_CHECK_ARG(N > 0 && T > 0 && H > 0 && W > 0, "bad shape");
  MSCL_CHECK_ARG(n_taps >= 1 && n_taps <= mscl::kMaxTaps && (n_taps & 1), "n_taps=%d must be odd and <= %d", n_taps,
                 mscl::kMaxTaps);
  MSCL_CHECK_ARG(n_taps / 2 < H && n_taps / 2 < W, "blur radius exceeds the frame (reflect border)");
  MSCL_CHECK_ARG(n_chunks > 0 && n_chunks <= 1024, "bad n_chunks");
  const int64_t THW = (int64_t)T * H * W;
  MSCL_CHECK_ARG(THW % 4 == 0 && (((uintptr_t)d_x) & 15) == 0, "T*H*W must be a multiple of 4 and x 16-byte aligned");
  const size_t smem = (size_t)2 * H * W * sizeof(float);
  MSCL_CHECK_ARG(smem <= 200 * 1024, "frame of %dx%d does not fit in shared memory", H, W);
  cudaStream_t s = mscl::as_stream(stream);
  MSCL_CHECK_ARG(!per_frame || ((int64_t)H * W) % 4 == 0, "H*W must be a multiple of 4 for per-frame parameters");
  MSCL_CHECK_ARG((int64_t)N * (per_frame ? T : 1) <= 65535, "too many clips / frames for one launch");
  const bool fast = n_taps == 11 && W <= 128 && W > 10 && H > 10 && W % 4 == 0;     // the config's case: 11 taps, 112x112 crops
  // bands of rows sized for ~3 CTAs per SM (<= 72 KB of staged rows incl. the 5 + 5 halo rows)
  int bands = 1;
  while (bands < H && (size_t)3 * ((H + bands - 1) / bands + 10) * W * 4 > 72 * 1024) ++bands;
  const int band_rows = (H + bands - 1) / bands;
  const size_t smem_fast = (size_t)3 * (band_rows + 10) * W * sizeof(float);
  const bool fast_ok = fast && band_rows >= 6 && smem_fast <= 200 * 1024;
  // per-frame parameters: the frame's mean luminance is formed inside the kernel by the cluster of its bands (no pre-pass)
  if (fast_ok && per_frame && bands <= 8 && (bands - 1) * band_rows < H) {
    MSCL_CUDA(mscl::ensure_dyn_smem(mscl::color_pipeline_fast_kernel<11, true>, smem_fast));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(N * T, bands);
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = smem_fast;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = bands;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MSCL_CUDA(cudaLaunchKernelEx(&cfg, mscl::color_pipeline_fast_kernel<11, true>, d_x, d_params, (const float *)d_gray_partial,
                                 (int)n_chunks, d_taps, d_norm, d_out, (int)T, (int)H, (int)W, band_rows, (int)per_frame));
    return MSCL_OK;
  }
  if (per_frame)
    mscl::clip_gray_sum_kernel<<<dim3(n_chunks, N * T), 256, 0, s>>>(d_x, d_gray_partial, THW, (int64_t)H * W, T);
  else
    mscl::clip_gray_sum_kernel<<<dim3(n_chunks, N), 256, 0, s>>>(d_x, d_gray_partial, THW, THW, 1);
  MSCL_LAUNCH_CHECK();
  if (fast_ok) {
    MSCL_CUDA(mscl::ensure_dyn_smem(mscl::color_pipeline_fast_kernel<11, false>, smem_fast));
    mscl::color_pipeline_fast_kernel<11, false><<<dim3(N * T, bands), 512, smem_fast, s>>>(d_x, d_params, d_gray_partial, n_chunks,
                                                                                           d_taps, d_norm, d_out, T, H, W, band_rows, per_frame);
    MSCL_LAUNCH_CHECK();
    return MSCL_OK;
  }
  MSCL_CUDA(mscl::ensure_dyn_smem(mscl::color_pipeline_kernel, smem));
  mscl::color_pipeline_kernel<<<N * T, 512, smem, s>>>(d_x, d_params, d_gray_partial, n_chunks, d_taps, n_taps, d_norm, d_out,
                                                       T, H, W, per_frame);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

}  // extern "C"
