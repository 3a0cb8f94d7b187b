// ABI bookkeeping: version, thread-local error text, device check.
#include <stdarg.h>
#include <string.h>

#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace mscl {

char *err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_err(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

cudaError_t ensure_dyn_smem_impl(const void *func, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void *, int>, size_t> configured;     // (kernel, device) -> bytes already allowed
  if (bytes <= 48 * 1024) return cudaSuccess;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  size_t &have = configured[std::make_pair(func, dev)];
  if (bytes <= have) return cudaSuccess;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) have = bytes;
  return e;
}

// One CTA copies a few KB from mapped pinned host memory into device memory (see mscl_fetch_host).
__global__ void __launch_bounds__(256)
fetch_host_kernel(const float4 *__restrict__ src, float4 *__restrict__ dst, int n4) {
  for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = src[i];
}

}  // namespace mscl

extern "C" {

int mscl_fetch_host(void *d_dst, const void *h_src_pinned, int64_t nbytes, mscl_stream_t stream) {
  MSCL_CHECK_ARG(d_dst && h_src_pinned, "null pointer");
  MSCL_CHECK_ARG(nbytes > 0 && nbytes % 16 == 0 && nbytes <= (1 << 20), "nbytes=%lld must be a multiple of 16 up to 1 MiB",
                 (long long)nbytes);
  MSCL_CHECK_ARG((((uintptr_t)d_dst | (uintptr_t)h_src_pinned) & 15) == 0, "pointers must be 16-byte aligned");
  void *src_dev = nullptr;
  cudaError_t e = cudaHostGetDevicePointer(&src_dev, const_cast<void *>(h_src_pinned), 0);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return mscl::set_err(MSCL_EINVAL, "h_src_pinned is not page-locked host memory mapped into the device: %s",
                         cudaGetErrorString(e));
  }
  mscl::fetch_host_kernel<<<1, 256, 0, mscl::as_stream(stream)>>>(reinterpret_cast<const float4 *>(src_dev),
                                                                   reinterpret_cast<float4 *>(d_dst), (int)(nbytes / 16));
  MSCL_LAUNCH_CHECK();
  return MSCL_OK;
}

int mscl_abi_version(void) { return MSCL_ABI_VERSION; }

const char *mscl_last_error(void) { return mscl::err_buf(); }

int mscl_device_check(int dev) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return mscl::set_err(MSCL_ECUDA, "no CUDA device: %s", cudaGetErrorString(e));
  }
  MSCL_CHECK_ARG(dev >= 0 && dev < n, "device %d out of range [0,%d)", dev, n);
  cudaDeviceProp p;
  MSCL_CUDA(cudaGetDeviceProperties(&p, dev));
  if (p.major != 10)
    return mscl::set_err(MSCL_EUNSUPPORTED,
                         "device %d is sm_%d%d; this library is built for sm_100a only",
                         dev, p.major, p.minor);
  return MSCL_OK;
}

}  // extern "C"
