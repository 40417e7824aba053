// api_common.cu -- vslam_last_error / version / device checks / pinned host memory of the C ABI.
#include "api_common.h"

#include <cstdarg>
#include <cstdio>

#include "../../include/vslam_b200.h"

namespace vslam {

static thread_local char g_error[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}

int require_device(int device) {
  int n = 0;
  const cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(VSLAM_ERR_CUDA, "no CUDA device: libvslam_b200 has no CPU fallback (%s)",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  if (device < 0 || device >= n) return fail(VSLAM_ERR_INVALID_ARGUMENT, "device %d outside [0, %d)", device, n);
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (major != 10) return fail(VSLAM_ERR_CUDA, "device %d has compute capability %d.x; this library holds sm_100a code only", device, major);
  CUDA_TRY(cudaSetDevice(device));
  return VSLAM_OK;
}

}  // namespace vslam

extern "C" {

const char* vslam_last_error(void) { return vslam::g_error; }
const char* vslam_version(void) { return "vslam_b200 0.1.0 sm_100a"; }

int vslam_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int vslam_host_alloc(void** ptr, size_t bytes) {
  if (!ptr) return vslam::fail(VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  // VSLAM_HOST_ALLOC_WC=1: write-combined pinned memory for buffers the host only WRITES (image upload sources): the
  // DMA engine reads it without snooping the CPU caches (measurement switch; CPU reads of such memory are slow)
  const char* e = std::getenv("VSLAM_HOST_ALLOC_WC");   // read per call: a host sets it around its upload buffers only
  const bool wc = e && atoi(e) != 0;
  CUDA_TRY(cudaHostAlloc(ptr, bytes ? bytes : 1, wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
  return VSLAM_OK;
}

int vslam_host_free(void* ptr) {
  if (ptr) CUDA_TRY(cudaFreeHost(ptr));
  return VSLAM_OK;
}

}  // extern "C"
