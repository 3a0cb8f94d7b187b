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

}  // namespace mscl

extern "C" {

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
