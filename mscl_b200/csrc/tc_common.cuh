// tcgen05 / TMA / mbarrier PTX wrappers and tensor-map builders shared by the tensor-core InfoNCE kernels
// (infonce_tc.cu: slab form, infonce_fused.cu: single-launch form).  sm_100a only.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace mscl {
namespace tc {

constexpr int kC = MSCL_DIM;          // 128 channels
constexpr int kLd = MSCL_PACK_LD;     // 132
constexpr int kRows = 128;            // query rows per CTA (UMMA M)

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- TMA ----
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, uint32_t src, int c0, int c1, int c2,
                                             int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void *src, uint32_t bytes,
                                             uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// ---- tcgen05 ----
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// One lane of a fully converged warp.  Issuing the tcgen05.mma stream from inside `if (elect_one())` in a
// branch the compiler can prove warp-uniform keeps descriptors in uniform registers (UIADD3 + UTCHMMA,
// 2 SASS instructions per dispatch); under a plain `lane == 0` branch every dispatch is wrapped in an
// ELECT/R2UR/BRA.U.ANY waterfall (~12 instructions), and the issuing thread, not the tensor pipe, sets the pace.
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred;
}
// D[tmem] (+)= A[tmem] . B[smem descriptor given as (lo, hi) words]
__device__ __forceinline__ void mma_ts_lh(uint32_t d_tmem, uint32_t a_tmem, uint32_t lo, uint32_t hi,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 bd;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 bd, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "r"(lo), "r"(hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                       uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem desc]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc,
                                       uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// smem matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 |
// version=1 <<46 | layout_type <<61  (2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type = 2) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
         ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | ((uint64_t)layout_type << 61);
}

#define TC_LD32(taddr, r)                                                                       \
  asm volatile(                                                                                 \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                 \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23," \
      "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),     \
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), \
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),           \
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),           \
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])            \
      : "r"(taddr))

#define TC_ST32(taddr, r)                                                                       \
  asm volatile(                                                                                 \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "                                          \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23," \
      "%24,%25,%26,%27,%28,%29,%30,%31};" ::"r"(r[0]),                                          \
      "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),   \
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),        \
      "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),       \
      "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),       \
      "r"(r[30]), "r"(r[31]), "r"(taddr)                                                        \
      : "memory")

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b),
               "f"(c), "f"(d)
               : "memory");
}

// ---- host side: tensor maps through the driver entry point (no link-time libcuda) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void *p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
    cudaGetLastError();
    return nullptr;
  }
  fn = (EncodeTiledFn)p;
  return fn;
}

// rows x 128 fp32 matrix with row pitch ld floats, viewed as {32, rows, 4} so that one box lands
// in shared memory as 4 channel-block slabs of [box_rows][128 B], 128-byte swizzled.
// A tensor map is a pure function of (pointer, rows, pitch, box, swizzle); the queues' maps never change and PyTorch's
// caching allocator hands the same q / positives blocks back step after step, so the ~14 driver encodes of a step launch
// (host time the GPU spends idle: the objective is launch-latency bound) are served from a small per-thread table.
struct MapKey {
  const void *ptr;
  int64_t rows;
  int ld, box_rows, swizzle;
  bool operator==(const MapKey &o) const {
    return ptr == o.ptr && rows == o.rows && ld == o.ld && box_rows == o.box_rows && swizzle == o.swizzle;
  }
};
struct MapCache {
  static constexpr int kSlots = 64;
  MapKey key[kSlots];
  CUtensorMap map[kSlots];
  bool used[kSlots];
  int next;
};

static int make_map(CUtensorMap *map, const float *ptr, int64_t rows, int ld, int box_rows,
                    CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  static thread_local MapCache cache = {};
  const MapKey k = {ptr, rows, ld, box_rows, (int)swizzle};
  for (int i = 0; i < MapCache::kSlots; ++i)
    if (cache.used[i] && cache.key[i] == k) {
      *map = cache.map[i];
      return MSCL_OK;
    }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_err(MSCL_EUNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {32, (cuuint64_t)rows, 4};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, 128};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 4};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(ptr), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_err(MSCL_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld ld=%d)",
                   (int)r, (long long)rows, ld);
  cache.key[cache.next] = k;
  cache.map[cache.next] = *map;
  cache.used[cache.next] = true;
  cache.next = (cache.next + 1) % MapCache::kSlots;
  return MSCL_OK;
}

// part [n_part][M][132] viewed as {32, M, 4, n_part}: one box = the O part of one CTA's slab
static int make_map_part(CUtensorMap *map, float *ptr, int M, int n_part) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_err(MSCL_EUNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[4] = {32, (cuuint64_t)M, 4, (cuuint64_t)n_part};
  cuuint64_t strides[3] = {(cuuint64_t)kLd * 4, 128, (cuuint64_t)M * kLd * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)kRows, 4, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, ptr, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_err(MSCL_ECUDA, "cuTensorMapEncodeTiled(part) failed with CUresult %d (M=%d n_part=%d)", (int)r, M,
                   n_part);
  return MSCL_OK;
}

}  // namespace tc
}  // namespace mscl
