// Shared helpers for the mscl_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mscl_b200.h"

namespace mscl {

// thread-local last-error text, returned by mscl_last_error()
char *err_buf();
int set_err(int code, const char *fmt, ...);

#define MSCL_CHECK_ARG(cond, ...)                           \
  do {                                                      \
    if (!(cond)) return ::mscl::set_err(MSCL_EINVAL, __VA_ARGS__); \
  } while (0)

#define MSCL_CUDA(call)                                                      \
  do {                                                                       \
    cudaError_t e__ = (call);                                                \
    if (e__ != cudaSuccess)                                                  \
      return ::mscl::set_err(MSCL_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, \
                             #call, cudaGetErrorString(e__));                \
  } while (0)

#define MSCL_LAUNCH_CHECK()                                                   \
  do {                                                                        \
    cudaError_t e__ = cudaGetLastError();                                     \
    if (e__ != cudaSuccess)                                                   \
      return ::mscl::set_err(MSCL_ECUDA, "%s:%d launch: %s", __FILE__,        \
                             __LINE__, cudaGetErrorString(e__));              \
  } while (0)

static inline cudaStream_t as_stream(mscl_stream_t s) { return (cudaStream_t)s; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is an attribute of (function, DEVICE): raise it to `bytes` on the current
// device unless this process already did so there (thread-safe; defined in abi.cu).  Returns a cudaError_t.
cudaError_t ensure_dyn_smem_impl(const void *func, size_t bytes);
template <typename F>
static inline cudaError_t ensure_dyn_smem(F func, size_t bytes) {
  return ensure_dyn_smem_impl(reinterpret_cast<const void *>(func), bytes);
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
// The reference hard-codes the decay base 0.99999 (moco.py:484) and evaluates
// `0.99999 ** float32_tensor` in float32, i.e. with the base rounded to
// float32(0.99999) = 0.9999899864196777.  log2 of THAT number:
constexpr float kLog2Decay = -1.4446615003766504e-05f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// round-to-nearest fp32 -> tf32 (10-bit mantissa); the tensor core itself truncates
__device__ __forceinline__ float to_tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float4 to_tf32_rn(float4 v) {
  return make_float4(to_tf32_rn(v.x), to_tf32_rn(v.y), to_tf32_rn(v.z), to_tf32_rn(v.w));
}

// ---- programmatic dependent launch (PDL) ----
// A kernel launched through launch_pdl may start while its predecessor in the stream is still
// running; it must execute pdl_wait() before touching anything the predecessor writes.  The
// predecessor calls pdl_trigger() as early as it likes (it only opens the gate for the successor's
// launch; memory visibility is pdl_wait's job, which blocks until the predecessor has completed).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// streaming 128-bit accesses (read-once / write-once data)
__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float4 *p, const float4 &v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p),
               "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

}  // namespace mscl
