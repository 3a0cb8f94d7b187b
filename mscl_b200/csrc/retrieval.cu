// K11: nearest-neighbour retrieval evaluation (tools/test_retrival.py:286-304 of the reference).
//
//   centre (subtract the column mean) -> L2-normalise rows -> sim = test @ train^T -> for k in (1,5,10,20,50):
//   acc@k = mean_i any(train_label[topk(sim_i, k)] == test_label_i)
//
// The GEMM is a plain library GEMM (cuBLAS through torch.matmul).  The kernels here do the two streaming ends:
//   * col_sum_partial / col_mean_finish / center_normalize : the centring + F.normalize(p=2, dim=1, eps=1e-12)
//   * retrieval_rank : "any of the top-k has my label"  <=>  fewer than k train items score above my best same-label
//     item.  One CTA per test row computes that rank (max over same-label columns, then a count), so the five
//     torch.topk calls + label gathers of the reference become one pass over sim; acc@k = mean(rank < k) for every k.
#include "common.cuh"

namespace mscl {

constexpr int kColThreads = 128;

// partial[chunk][d] = sum of x[r][d] over the rows of the chunk, accumulated in double (fixed order: deterministic)
__global__ void __launch_bounds__(kColThreads)
col_sum_partial_kernel(const float *__restrict__ x, int64_t N, int D, int rows_per_chunk, double *__restrict__ partial) {
  const int d = blockIdx.x * kColThreads + threadIdx.x;
  if (d >= D) return;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
  const int64_t r1 = min(N, r0 + rows_per_chunk);
  double acc = 0.0;
  for (int64_t r = r0; r < r1; ++r) acc += (double)__ldg(x + r * D + d);
  partial[(int64_t)blockIdx.y * D + d] = acc;
}

__global__ void __launch_bounds__(kColThreads)
col_mean_finish_kernel(const double *__restrict__ partial, int n_chunks, int D, int64_t N, float *__restrict__ mean) {
  const int d = blockIdx.x * kColThreads + threadIdx.x;
  if (d >= D) return;
  double acc = 0.0;
  for (int c = 0; c < n_chunks; ++c) acc += partial[(int64_t)c * D + d];
  mean[d] = (float)(acc / (double)N);
}

// one warp per row: y = (x - mean) / max(||x - mean||_2, 1e-12)
__global__ void __launch_bounds__(256)
center_normalize_kernel(const float *__restrict__ x, const float *__restrict__ mean, int64_t N, int D, float *__restrict__ y) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  const int lane = threadIdx.x & 31;
  const float *xr = x + row * D;
  float ss = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float v = __fsub_rn(xr[d], __ldg(mean + d));
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  const float denom = fmaxf(sqrtf(ss), 1e-12f);
  float *yr = y + row * D;
  for (int d = lane; d < D; d += 32) yr[d] = __fdiv_rn(__fsub_rn(xr[d], __ldg(mean + d)), denom);
}

constexpr int kRankThreads = 256;

// rank[i] = #{ j : sim[i][j] > max_{j' : train_label[j'] == test_label[i]} sim[i][j'] }   (n_train if no such j')
__global__ void __launch_bounds__(kRankThreads)
retrieval_rank_kernel(const float *__restrict__ sim, int64_t ld, const int64_t *__restrict__ train_label,
                      const int64_t *__restrict__ test_label, int n_train, int32_t *__restrict__ rank) {
  __shared__ float s_max[kRankThreads / 32];
  __shared__ int s_cnt[kRankThreads / 32];
  __shared__ float s_best;
  const int i = blockIdx.x;
  const float *row = sim + (int64_t)i * ld;
  const int64_t mine = test_label[i];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float best = -INFINITY;
  for (int j = threadIdx.x; j < n_train; j += kRankThreads)
    if (__ldg(train_label + j) == mine) best = fmaxf(best, row[j]);
  best = warp_max(best);
  if (lane == 0) s_max[warp] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    float b = s_max[0];
    for (int w = 1; w < kRankThreads / 32; ++w) b = fmaxf(b, s_max[w]);
    s_best = b;
  }
  __syncthreads();
  best = s_best;
  int cnt = 0;
  for (int j = threadIdx.x; j < n_train; j += kRankThreads) cnt += (row[j] > best) ? 1 : 0;   // second read: L1/L2
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) s_cnt[warp] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int c = 0;
    for (int w = 0; w < kRankThreads / 32; ++w) c += s_cnt[w];
    rank[i] = (best == -INFINITY) ? n_train : c;
  }
}

}  // namespace mscl

extern "C" int mscl_center_normalize(const float *d_x, int64_t N, int32_t D, double *d_partial, int32_t n_chunks,
                                     float *d_mean, float *d_out, mscl_stream_t stream) {
  MSCL_CHECK_ARG(d_x && d_partial && d_mean && d_out, "null pointer");
  MSCL_CHECK_ARG(N > 0 && D > 0 && n_chunks > 0 && n_chunks <= 65535, "bad N=%lld D=%d n_chunks=%d", (long long)N, D, n_chunks);
  cudaStream_t s = mscl::as_stream(stream);
  const int rows_per_chunk = (int)((N + n_chunks - 1) / n_chunks);
  const int col_blocks = (D + mscl::kColThreads - 1) / mscl::kColThreads;
  mscl::col_sum_partial_kernel<<<dim3(col_blocks, n_chunks), mscl::kColThreads, 0, s>>>(d_x, N, D, rows_per_chunk, d_partial);
  MSCL_LAUNCH_CHECK();
  mscl::col_mean_finish_kernel<<<col_blocks, mscl::kColThreads, 0, s>>>(d_partial, n_chunks, D, N, d_mean);
  MSCL_LAUNCH_CHECK();
  const int rows_per_cta = 256 / 32;
  mscl::center_normalize_kernel<<<(unsigned)((N + rows_per_cta - 1) / rows_per_cta), 256, 0, s>>>(d_x, d_mean, N, D, d_out);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

extern "C" int mscl_retrieval_rank(const float *d_sim, int64_t ld, const int64_t *d_train_label, const int64_t *d_test_label,
                                   int32_t n_test, int32_t n_train, int32_t *d_rank, mscl_stream_t stream) {
  MSCL_CHECK_ARG(d_sim && d_train_label && d_test_label && d_rank, "null pointer");
  MSCL_CHECK_ARG(n_test > 0 && n_train > 0 && ld >= n_train, "bad n_test=%d n_train=%d ld=%lld", n_test, n_train, (long long)ld);
  mscl::retrieval_rank_kernel<<<n_test, mscl::kRankThreads, 0, mscl::as_stream(stream)>>>(d_sim, ld, d_train_label,
                                                                                          d_test_label, n_train, d_rank);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}
