// K5: ring-buffer enqueue of the negative queue, plus the layout converters that
// materialise the reference's (C, K) `queue`, int64 `count` and decayed `weight`
// buffers on demand, plus the K6 row gather used by shuffle-BN.
//
// Device layout is key-major: queue[K_local][C] so that an enqueue is one
// contiguous, fully coalesced block write of B_all*C floats and a queue tile is a
// K-major tcgen05 operand.  Ages are implicit: count[j] = n_enq - birth[j], so the
// reference's K-long `count += 1` read-modify-write (moco.py:427) disappears.
#include "common.cuh"

namespace mscl {

// One thread per float4 of the key block.
__global__ void __launch_bounds__(256)
enqueue_kernel(float *__restrict__ queue, float *__restrict__ queue_tf32, int32_t *__restrict__ birth,
               int64_t *__restrict__ qstate, const float *__restrict__ keys,
               int B_all, int C4, int64_t K_total, int64_t shard_begin,
               int64_t K_local, float *__restrict__ saved,
               int32_t *__restrict__ saved_birth) {
  const int64_t ptr = *reinterpret_cast<volatile int64_t *>(&qstate[0]);
  const int64_t n_enq = *reinterpret_cast<volatile int64_t *>(&qstate[1]);
  const int64_t total = (int64_t)B_all * C4;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total;
       v += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(v / C4);
    const int c4 = (int)(v - (int64_t)i * C4);
    int64_t slot = ptr + i;
    if (slot >= K_total) slot -= K_total;
    const int64_t local = slot - shard_begin;
    if (local < 0 || local >= K_local) continue;
    float4 *dst = reinterpret_cast<float4 *>(queue + local * (int64_t)C4 * 4) + c4;
    if (saved != nullptr) {
      // the overwritten row, in the form the tensor-core pass reads (the rounded operand copy when there is one): what
      // mscl_infonce_fused_multi_x streams for the rows that still read the queue as it was before this enqueue
      reinterpret_cast<float4 *>(saved)[v] =
          queue_tf32 != nullptr ? reinterpret_cast<float4 *>(queue_tf32 + local * (int64_t)C4 * 4)[c4] : *dst;
      if (c4 == 0) saved_birth[i] = birth[local];
    }
    const float4 kv = __ldg(reinterpret_cast<const float4 *>(keys) + v);
    *dst = kv;
    if (queue_tf32 != nullptr)
      reinterpret_cast<float4 *>(queue_tf32 + local * (int64_t)C4 * 4)[c4] = to_tf32_rn(kv);
    if (c4 == 0) birth[local] = (int32_t)n_enq;  // age becomes 1 after n_enq+1
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned long long done =
        atomicAdd(reinterpret_cast<unsigned long long *>(&qstate[2]), 1ULL);
    if (done == (unsigned long long)gridDim.x - 1ULL) {
      qstate[0] = (ptr + B_all) % K_total;
      qstate[1] = n_enq + 1;
      qstate[2] = 0;
      __threadfence();
    }
  }
}

// 32x32 smem transpose between key-major [K][C] and channel-major [C][K].
// mode 0: export raw; mode 1: export decayed weight; mode 2: import.
template <int MODE>
__global__ void __launch_bounds__(256)
queue_transpose_kernel(float *__restrict__ queue_kc, float *__restrict__ queue_tf32,
                       int32_t *__restrict__ birth,
                       const int64_t *__restrict__ qstate, float *__restrict__ ck,
                       int64_t *__restrict__ count, int C, int64_t K_local) {
  __shared__ float tile[32][33];
  const int64_t n_enq = qstate[1];
  const int64_t k0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  if (MODE == 2) {
    // read [C][K] rows (K contiguous), write [K][C]
    for (int r = ty; r < 32; r += 8) {
      const int c = c0 + r;
      const int64_t k = k0 + tx;
      tile[r][tx] = (c < C && k < K_local) ? ck[(int64_t)c * K_local + k] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
      const int64_t k = k0 + r;
      const int c = c0 + tx;
      if (k < K_local && c < C) {
        queue_kc[k * C + c] = tile[tx][r];
        if (queue_tf32 != nullptr) queue_tf32[k * C + c] = to_tf32_rn(tile[tx][r]);
      }
    }
    if (blockIdx.y == 0 && threadIdx.x < 32) {
      const int64_t k = k0 + threadIdx.x;
      if (k < K_local) birth[k] = (int32_t)(n_enq - count[k]);
    }
  } else {
    for (int r = ty; r < 32; r += 8) {
      const int64_t k = k0 + r;
      const int c = c0 + tx;
      float v = 0.f;
      if (k < K_local && c < C) {
        v = queue_kc[k * C + c];
        if (MODE == 1) {
          // reference: 0.99999 ** (1.0 * count) in float32, then mul (moco.py:484-485)
          const float age = (float)(n_enq - (int64_t)birth[k]);
          v = __fmul_rn(v, powf(0.99999f, age));
        }
      }
      tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
      const int c = c0 + r;
      const int64_t k = k0 + tx;
      if (c < C && k < K_local) ck[(int64_t)c * K_local + k] = tile[tx][r];
    }
    if (MODE == 0 && count != nullptr && blockIdx.y == 0 && threadIdx.x < 32) {
      const int64_t k = k0 + threadIdx.x;
      if (k < K_local) count[k] = n_enq - (int64_t)birth[k];
    }
  }
}

__global__ void __launch_bounds__(256)
gather_rows_kernel(const float4 *__restrict__ x, const int64_t *__restrict__ idx,
                   float4 *__restrict__ out, int64_t row_vec) {
  const int r = blockIdx.y;
  const int64_t src = idx[r];
  const float4 *s = x + src * row_vec;
  float4 *d = out + (int64_t)r * row_vec;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x * 4 + threadIdx.x; v < row_vec;
       v += (int64_t)gridDim.x * blockDim.x * 4) {
    float4 a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (v + u * blockDim.x < row_vec) a[u] = ldg_stream(s + v + u * blockDim.x);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (v + u * blockDim.x < row_vec) stg_stream(d + v + u * blockDim.x, a[u]);
  }
}

}  // namespace mscl

extern "C" {

int mscl_enqueue(float *d_queue, float *d_queue_tf32, int32_t *d_birth, int64_t *d_qstate,
                 const float *d_keys, int32_t B_all, int32_t C, int64_t K_total,
                 int64_t shard_begin, int64_t K_local, float *d_saved,
                 int32_t *d_saved_birth, mscl_stream_t stream) {
  MSCL_CHECK_ARG(d_queue && d_birth && d_qstate && d_keys, "null pointer");
  MSCL_CHECK_ARG(B_all > 0 && C > 0 && C % 4 == 0, "bad B_all=%d C=%d", B_all, C);
  MSCL_CHECK_ARG(K_total > 0 && K_total % B_all == 0,
                 "K=%lld must be a multiple of the gathered batch %d (moco.py:432)",
                 (long long)K_total, B_all);
  MSCL_CHECK_ARG(shard_begin >= 0 && K_local > 0 && shard_begin + K_local <= K_total,
                 "bad shard [%lld,+%lld) of %lld", (long long)shard_begin,
                 (long long)K_local, (long long)K_total);
  MSCL_CHECK_ARG((d_saved == nullptr) == (d_saved_birth == nullptr),
                 "d_saved and d_saved_birth must both be given or both be NULL");
  const int64_t total = (int64_t)B_all * (C / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 1184) blocks = 1184;  // 8 CTAs x 148 SMs
  mscl::enqueue_kernel<<<blocks, 256, 0, mscl::as_stream(stream)>>>(
      d_queue, d_queue_tf32, d_birth, d_qstate, d_keys, B_all, C / 4, K_total, shard_begin,
      K_local, d_saved, d_saved_birth);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

static int transpose_launch(int mode, float *q, float *q_tf32, int32_t *birth, const int64_t *qstate,
                            float *ck, int64_t *count, int32_t C, int64_t K_local,
                            mscl_stream_t stream) {
  MSCL_CHECK_ARG(q && birth && qstate && ck, "null pointer");
  MSCL_CHECK_ARG(C > 0 && K_local > 0, "bad C=%d K_local=%lld", C, (long long)K_local);
  dim3 grid((unsigned)((K_local + 31) / 32), (unsigned)((C + 31) / 32));
  cudaStream_t s = mscl::as_stream(stream);
  if (mode == 0)
    mscl::queue_transpose_kernel<0><<<grid, 256, 0, s>>>(q, q_tf32, birth, qstate, ck, count, C, K_local);
  else if (mode == 1)
    mscl::queue_transpose_kernel<1><<<grid, 256, 0, s>>>(q, q_tf32, birth, qstate, ck, count, C, K_local);
  else
    mscl::queue_transpose_kernel<2><<<grid, 256, 0, s>>>(q, q_tf32, birth, qstate, ck, count, C, K_local);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

int mscl_queue_export(const float *d_queue, const int32_t *d_birth,
                      const int64_t *d_qstate, float *d_queue_ck, int64_t *d_count,
                      int32_t C, int64_t K_local, mscl_stream_t stream) {
  MSCL_CHECK_ARG(d_count != nullptr, "null count");
  return transpose_launch(0, const_cast<float *>(d_queue), nullptr, const_cast<int32_t *>(d_birth),
                          d_qstate, d_queue_ck, d_count, C, K_local, stream);
}

int mscl_queue_import(float *d_queue, float *d_queue_tf32, int32_t *d_birth, const int64_t *d_qstate,
                      const float *d_queue_ck, const int64_t *d_count, int32_t C,
                      int64_t K_local, mscl_stream_t stream) {
  MSCL_CHECK_ARG(d_count != nullptr, "null count");
  return transpose_launch(2, d_queue, d_queue_tf32, d_birth, d_qstate, const_cast<float *>(d_queue_ck),
                          const_cast<int64_t *>(d_count), C, K_local, stream);
}

int mscl_queue_weight(const float *d_queue, const int32_t *d_birth,
                      const int64_t *d_qstate, float *d_weight_ck, int32_t C,
                      int64_t K_local, mscl_stream_t stream) {
  return transpose_launch(1, const_cast<float *>(d_queue), nullptr, const_cast<int32_t *>(d_birth),
                          d_qstate, d_weight_ck, nullptr, C, K_local, stream);
}

int mscl_gather_rows(const float *d_x, const int64_t *d_idx, float *d_out,
                     int32_t n_rows, int64_t row_elems, mscl_stream_t stream) {
  MSCL_CHECK_ARG(d_x && d_idx && d_out, "null pointer");
  MSCL_CHECK_ARG(n_rows > 0 && row_elems > 0 && row_elems % 4 == 0,
                 "bad n_rows=%d row_elems=%lld", n_rows, (long long)row_elems);
  MSCL_CHECK_ARG(((uintptr_t)d_x % 16) == 0 && ((uintptr_t)d_out % 16) == 0,
                 "pointers must be 16-byte aligned");
  const int64_t row_vec = row_elems / 4;
  int bx = (int)((row_vec + 1023) / 1024);
  if (bx < 1) bx = 1;
  if (bx > 4096) bx = 4096;
  dim3 grid((unsigned)bx, (unsigned)n_rows);
  mscl::gather_rows_kernel<<<grid, 256, 0, mscl::as_stream(stream)>>>(
      reinterpret_cast<const float4 *>(d_x), d_idx, reinterpret_cast<float4 *>(d_out),
      row_vec);
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

}  // extern "C"
