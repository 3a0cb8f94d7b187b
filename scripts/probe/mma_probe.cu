// Stand-alone probe of the tcgen05 operand forms used by the InfoNCE kernel.
//   mode 0: SS, A K-major smem, B K-major smem       D[128x64]  = Q  . W^T
//   mode 1: TS, A in TMEM,      B K-major smem       D[128x64]  = Q  . W^T
//   mode 2: SS, A K-major smem, B MN-major smem      D[128x128] = P  . W
//   mode 3: TS, A in TMEM,      B MN-major smem      D[128x128] = P  . W
// Q [128x128], W [64 keys x 128 ch], P [128 x 64] (row-major fp32 in global memory).
#include <vector>
#include <cstdlib>
#include <cmath>
#include "../../mscl_b200/csrc/infonce_tc.cu"

using namespace mscl::tc;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tw,
             const __grid_constant__ CUtensorMap tp, const __grid_constant__ CUtensorMap tw2, const float *Q, const float *P, float *D, int mode,
             uint32_t lbo, uint32_t sbo, uint32_t lt) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base, sW = base + 65536, sP = base + 65536 + 32768;
  const uint32_t sW2 = base + 65536 + 32768 + 32768;
  const uint32_t bar = base + 65536 + 32768 + 32768 + 32768, bar2 = bar + 8;
  volatile uint32_t *tptr = reinterpret_cast<volatile uint32_t *>(gbase + 65536 + 32768 + 32768 + 32768 + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + 65536 + 32768 + 32768 + 32768 + 16), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 65536 + 32768 + 32768 + 32768);
    tma_load_3d(sW2, &tw2, bar, 0, 0, 0);
    tma_load_3d(sQ, &tq, bar, 0, 0, 0);
    tma_load_3d(sW, &tw, bar, 0, 0, 0);
    tma_load_3d(sP, &tp, bar, 0, 0, 0);
  }
  // A into TMEM at columns 256.. : row = thread
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  const int r = threadIdx.x;
  const float *src = (mode == 1) ? Q + r * 128 : P + r * 64;
  const int ncol = (mode == 1) ? 128 : 64;
  for (int h = 0; h < ncol / 32; ++h) {
    uint32_t v[32];
    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(src[h * 32 + j]);
    TC_ST32(lane_base + 256 + h * 32, v);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    mbar_wait(bar, 0);
    tc_fence_after();
    if (mode == 0 || mode == 1) {
      for (int cb = 0; cb < 4; ++cb)
        for (int ks = 0; ks < 4; ++ks) {
          // lt == 1: read the K-major operand from the 32B-atom-swizzled copy (could one copy serve both MMAs?)
          const uint64_t bd = lt == 1 ? make_desc(sW2 + cb * 8192 + ks * 32, lbo, sbo, 1) : make_desc(sW + cb * 8192 + ks * 32, 16, 1024);
          if (mode == 0)
            mma_ss(tmem, make_desc(sQ + cb * 16384 + ks * 32, 16, 1024), bd, kIdesc1, (cb | ks) ? 1u : 0u);
          else
            mma_ts(tmem, tmem + 256 + cb * 32 + ks * 8, bd, kIdesc1, (cb | ks) ? 1u : 0u);
        }
    } else {
      for (int j = 0; j < 8; ++j) {
        uint64_t bd = make_desc((lt == 1 ? sW2 : sW) + j * 1024, lbo, sbo);
        if (lt == 1) bd = (bd & ~(7ull << 61)) | (1ull << 61);
        if (mode == 2)   // P in smem K-major: 2 channel-block slabs of [128 rows][128 B]; k-step j -> slab j/4, 32B chunk j%4
          mma_ss(tmem, make_desc(sP + (j / 4) * 16384 + (j % 4) * 32, 16, 1024), bd, kIdesc2, j ? 1u : 0u);
        else
          mma_ts(tmem, tmem + 256 + j * 8, bd, kIdesc2, j ? 1u : 0u);
      }
    }
    tc_commit(bar2);
  }
  mbar_wait(bar2, 0);
  tc_fence_after();
  const int nout = (mode < 2) ? 64 : 128;
  for (int h = 0; h < nout / 32; ++h) {
    uint32_t v[32];
    TC_LD32(lane_base + h * 32, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) D[r * 128 + h * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

static int make_map_p(CUtensorMap *map, const float *ptr) {   // P [128 x 64] -> {32, 128, 2}
  EncodeTiledFn fn = get_encode_fn();
  cuuint64_t dims[3] = {32, 128, 2};
  cuuint64_t strides[2] = {64 * 4, 128};
  cuuint32_t box[3] = {32, 128, 2};
  cuuint32_t estr[3] = {1, 1, 1};
  return (int)fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(ptr), dims, strides, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

static float tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y; }

int main() {
  std::vector<float> Q(128 * 128), W(64 * 128), P(128 * 64), D(128 * 128);
  srand(1);
  for (auto &x : Q) x = tf32((rand() % 2001 - 1000) / 1000.f);
  for (auto &x : W) x = tf32((rand() % 2001 - 1000) / 1000.f);
  for (auto &x : P) x = tf32((rand() % 2001 - 1000) / 1000.f);
  float *dQ, *dW, *dP, *dD;
  cudaMalloc(&dQ, Q.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dP, P.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dQ, Q.data(), Q.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dP, P.data(), P.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap tq, tw, tp, tw2;
  {
    EncodeTiledFn fn = get_encode_fn();
    cuuint64_t dims[3] = {32, 64, 4}; cuuint64_t strides[2] = {512, 128}; cuuint32_t box[3] = {32, 64, 4}; cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(&tw2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dW, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("tw2 encode failed %d\n", (int)r); return 1; }
  }
  if (make_map(&tq, dQ, 128, 128, 128) || make_map(&tw, dW, 64, 128, 64) || make_map_p(&tp, dP)) { printf("map fail %s\n", mscl_last_error()); return 1; }
  const int smem = 65536 + 32768 + 32768 + 32768 + 64 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  struct Cfg { int mode; uint32_t lbo, sbo; const char *name; uint32_t lt; };
  Cfg cfgs[] = {{0, 16, 1024, "SS K/K"}, {1, 16, 1024, "TS K"}, {2, 8192, 1024, "SS MN lbo=8192 sbo=1024"}, {3, 8192, 1024, "TS MN lbo=8192 sbo=1024"},
                {2, 1024, 8192, "SS MN lbo=1024 sbo=8192"}, {3, 1024, 8192, "TS MN lbo=1024 sbo=8192"},
                {2, 8192, 512, "SS MN32 lbo=8192 sbo=512", 1}, {3, 8192, 512, "TS MN32 lbo=8192 sbo=512", 1},
                {2, 512, 8192, "SS MN32 lbo=512 sbo=8192", 1}, {2, 8192, 1024, "SS MN32 lbo=8192 sbo=1024", 1},
                {1, 16, 1024, "TS K from 32B-atom copy sbo=1024", 1}, {0, 16, 1024, "SS K from 32B-atom copy sbo=1024", 1},
                {1, 16, 512, "TS K from 32B-atom copy sbo=512", 1}, {1, 16, 256, "TS K from 32B-atom copy sbo=256", 1}};
  for (auto &c : cfgs) {
    cudaMemset(dD, 0xff, D.size() * 4);
    probe_kernel<<<1, 128, smem>>>(tq, tw, tp, tw2, dQ, dP, dD, c.mode, c.lbo, c.sbo, c.lt);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-28s CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0; int nz = 0;
    const int nout = c.mode < 2 ? 64 : 128;
    for (int i = 0; i < 128; ++i)
      for (int n = 0; n < nout; ++n) {
        double ref = 0;
        if (c.mode < 2) for (int k = 0; k < 128; ++k) ref += (double)Q[i * 128 + k] * W[n * 128 + k];
        else for (int k = 0; k < 64; ++k) ref += (double)P[i * 64 + k] * W[k * 128 + n];
        double got = D[i * 128 + n];
        if (got != 0) nz++;
        maxerr = fmax(maxerr, fabs(got - ref)); maxref = fmax(maxref, fabs(ref));
      }
    printf("%-28s max|err| %.3e  max|ref| %.3e  nonzero %d/%d  D[0][0..3]= %g %g %g %g\n", c.name, maxerr, maxref, nz, 128 * nout,
           D[0], D[1], D[2], D[3]);
  }
  return 0;
}
