// Timing probe: cycles per tcgen05.mma dispatch for the operand forms the InfoNCE kernel could use.
// Operand contents are whatever shared memory / TMEM hold (zero-filled); only the pacing matters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr \
//        -o scripts/probe/mma_rate scripts/probe/mma_rate.cu mscl_b200/csrc/abi.cu
#include <vector>
#include <cstdlib>
#include <cmath>
#include "../../mscl_b200/csrc/infonce_tc.cu"

using namespace mscl::tc;

__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// descriptor passed as (lo, hi) words: lo = base + constant, no dependent chain per dispatch
__device__ __forceinline__ void mma_ts_lh(uint32_t d, uint32_t a, uint32_t lo, uint32_t hi, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 bd, {%2, %3};\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n\t}\n" ::"r"(d), "r"(a), "r"(lo), "r"(hi), "r"(idesc), "r"(acc) : "memory");
}

// idesc for kind::f16 with bf16 operands: c_format f32 (1<<4), a_format bf16 (1<<7), b_format bf16 (1<<10)
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(128, 1) rate_kernel(int mode, int rounds, long long *out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sA = base, sB = base + 65536, sB2 = base + 65536 + 65536;
  const uint32_t bar = base + 3 * 65536, tp = bar + 16;
  volatile uint32_t *tptr = reinterpret_cast<volatile uint32_t *>(gbase + 3 * 65536 + 16);
  for (int i = threadIdx.x; i < 3 * 65536 / 4; i += 128) reinterpret_cast<uint32_t *>(gbase)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tp), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      switch (mode) {
        case 0:  // tf32 MMA1 TS: 16 x [128x64x8], A TMEM, B K-major SW128 (current kernel)
          for (int cb = 0; cb < 4; ++cb)
            for (int ks = 0; ks < 4; ++ks)
              mma_ts(tmem + 128, tmem + 256 + cb * 32 + ks * 8, make_desc(sB + cb * 8192 + ks * 32, 16, 1024), idesc_tf32(128, 64, 0), 1);
          break;
        case 1:  // tf32 MMA1 SS
          for (int cb = 0; cb < 4; ++cb)
            for (int ks = 0; ks < 4; ++ks)
              mma_ss(tmem + 128, make_desc(sA + cb * 16384 + ks * 32, 16, 1024), make_desc(sB + cb * 8192 + ks * 32, 16, 1024),
                     idesc_tf32(128, 64, 0), 1);
          break;
        case 2:  // tf32 MMA2 TS, B MN-major 32B-atom swizzle: 8 x [128x128x8] (current kernel)
          for (int j = 0; j < 8; ++j)
            mma_ts(tmem, tmem + 128 + j * 8, make_desc(sB2 + j * 1024, 8192, 512, 1), idesc_tf32(128, 128, 1), 1);
          break;
        case 3:  // tf32 MMA2 TS, B K-major SW128 (as if a transposed tile [128 ch][64 keys] existed): 8 x [128x128x8]
          for (int j = 0; j < 8; ++j)
            mma_ts(tmem, tmem + 128 + j * 8, make_desc(sB2 + (j / 4) * 16384 + (j % 4) * 32, 16, 1024), idesc_tf32(128, 128, 0), 1);
          break;
        case 4:  // bf16 MMA1 TS: 8 x [128x64x16], B K-major SW128 (tile [64 keys][128 ch] bf16 = 2 slabs of 64 ch)
          for (int j = 0; j < 8; ++j)
            mma_f16_ts(tmem + 128, tmem + 256 + j * 8, make_desc(sB + (j / 4) * 8192 + (j % 4) * 32, 16, 1024), idesc_bf16(128, 64, 0), 1);
          break;
        case 5:  // bf16 MMA2 TS, B MN-major SW128 (tile [64 keys][128 ch] bf16, ch contiguous): 4 x [128x128x16]
          for (int j = 0; j < 4; ++j)
            mma_f16_ts(tmem, tmem + 128 + j * 8, make_desc(sB2 + j * 2048, 8192, 1024), idesc_bf16(128, 128, 1), 1);
          break;
        case 6:  // bf16 MMA1 SS
          for (int j = 0; j < 8; ++j)
            mma_f16_ss(tmem + 128, make_desc(sA + (j / 4) * 16384 + (j % 4) * 32, 16, 1024),
                       make_desc(sB + (j / 4) * 8192 + (j % 4) * 32, 16, 1024), idesc_bf16(128, 64, 0), 1);
          break;
        case 7:  // tf32 MMA1 TS with N=128 keys per dispatch (8 x... 16 x [128x128x8])
          for (int cb = 0; cb < 4; ++cb)
            for (int ks = 0; ks < 4; ++ks)
              mma_ts(tmem, tmem + 256 + cb * 32 + ks * 8, make_desc(sB + cb * 16384 + ks * 32, 16, 1024), idesc_tf32(128, 128, 0), 1);
          break;
        case 8:  // tf32 MMA1 TS N=256
          for (int cb = 0; cb < 4; ++cb)
            for (int ks = 0; ks < 4; ++ks)
              mma_ts(tmem, tmem + 256 + cb * 32 + ks * 8, make_desc(sB + (cb & 1) * 32768 + ks * 32, 16, 1024), idesc_tf32(128, 256, 0), 1);
          break;
        case 9:  // tf32 MMA2 SS, B MN-major 32B atom, A = P in smem K-major
          for (int j = 0; j < 8; ++j)
            mma_ss(tmem, make_desc(sA + (j / 4) * 16384 + (j % 4) * 32, 16, 1024), make_desc(sB2 + j * 1024, 8192, 512, 1),
                   idesc_tf32(128, 128, 1), 1);
          break;
        case 11: {  // tf32 MMA1 TS 16x[128x64x8], descriptors as base_lo + immediate
          const uint64_t d0 = make_desc(sB + (r % 3) * 32768, 16, 1024);
          const uint32_t lo = (uint32_t)d0, hi = (uint32_t)(d0 >> 32);
#pragma unroll
          for (int cb = 0; cb < 4; ++cb)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              mma_ts_lh(tmem + 128, tmem + 256 + cb * 32 + ks * 8, lo + ((cb * 8192 + ks * 32) >> 4), hi, idesc_tf32(128, 64, 0), 1);
          break;
        }
        case 12: {  // tf32 MMA2 TS 8x[128x128x8] B MN 32B-atom, descriptors as base_lo + immediate
          const uint64_t d0 = make_desc(sB2 + (r % 2) * 32768, 8192, 512, 1);
          const uint32_t lo = (uint32_t)d0, hi = (uint32_t)(d0 >> 32);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            mma_ts_lh(tmem, tmem + 128 + j * 8, lo + ((j * 1024) >> 4), hi, idesc_tf32(128, 128, 1), 1);
          break;
        }
        case 13: {  // the kernel's alternation: MMA1 (16 x N=64) then MMA2 (8 x N=128), lo/hi descriptors
          const uint64_t d0 = make_desc(sB + (r % 2) * 32768, 16, 1024);
          const uint32_t lo = (uint32_t)d0, hi = (uint32_t)(d0 >> 32);
#pragma unroll
          for (int cb = 0; cb < 4; ++cb)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              mma_ts_lh(tmem + 128 + (r & 1) * 64, tmem + 256 + cb * 32 + ks * 8, lo + ((cb * 8192 + ks * 32) >> 4), hi, idesc_tf32(128, 64, 0), (cb | ks) ? 1u : 0u);
          const uint64_t e0 = make_desc(sB2 + (r % 2) * 32768, 8192, 512, 1);
          const uint32_t lo2 = (uint32_t)e0, hi2 = (uint32_t)(e0 >> 32);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            mma_ts_lh(tmem, tmem + 128 + ((r + 1) & 1) * 64 + j * 8, lo2 + ((j * 1024) >> 4), hi2, idesc_tf32(128, 128, 1), 1);
          break;
        }
        case 10:  // bf16 MMA2 TS with B K-major
          for (int j = 0; j < 4; ++j)
            mma_f16_ts(tmem, tmem + 128 + j * 8, make_desc(sB2 + j * 32, 16, 1024), idesc_bf16(128, 128, 0), 1);
          break;
      }
    }
    tc_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  const int smem = 3 * 65536 + 64 + 1024;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long *d_out;
  cudaMalloc(&d_out, 148 * 8);
  const char *names[] = {"tf32 MMA1 TS  16x[128x64x8]  B K-major", "tf32 MMA1 SS  16x[128x64x8]", "tf32 MMA2 TS  8x[128x128x8] B MN 32B-atom",
                         "tf32 MMA2 TS  8x[128x128x8] B K-major", "bf16 MMA1 TS  8x[128x64x16]", "bf16 MMA2 TS  4x[128x128x16] B MN-major",
                         "bf16 MMA1 SS  8x[128x64x16]", "tf32 MMA1 TS 16x[128x128x8]", "tf32 MMA1 TS 16x[128x256x8]",
                         "tf32 MMA2 SS  8x[128x128x8] B MN 32B-atom", "bf16 MMA2 TS  4x[128x128x16] B K-major",
                         "tf32 MMA1 TS 16x[128x64x8] lo/hi desc", "tf32 MMA2 TS 8x[128x128x8] MN32 lo/hi desc", "tf32 MMA1+MMA2 alternating lo/hi"};
  const double macs[] = {128. * 64 * 128, 128. * 64 * 128, 128. * 128 * 64, 128. * 128 * 64, 128. * 64 * 128, 128. * 128 * 64, 128. * 64 * 128,
                         128. * 128 * 128, 128. * 256 * 128, 128. * 128 * 64, 128. * 128 * 64,
                         128. * 64 * 128, 128. * 128 * 64, 2 * 128. * 64 * 128};
  for (int grid : {148}) {
    for (int mode = 0; mode < 14; ++mode) {
      const int rounds = 200;
      rate_kernel<<<grid, 128, smem>>>(mode, rounds, d_out);
      rate_kernel<<<grid, 128, smem>>>(mode, rounds, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
      std::vector<long long> h(grid);
      cudaMemcpy(h.data(), d_out, grid * 8, cudaMemcpyDeviceToHost);
      double avg = 0;
      for (auto v : h) avg += (double)v / grid;
      printf("grid %3d  %-44s %8.1f cyc/round  %6.1f MAC/cyc/SM\n", grid, names[mode], avg / rounds, macs[mode] * rounds / avg);
    }
  }
  return 0;
}
