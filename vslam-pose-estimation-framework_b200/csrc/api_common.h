// api_common.h -- error plumbing shared by the C ABI translation units.
#pragma once

#include <cuda_runtime.h>

namespace vslam {

// records the message for vslam_last_error() on this thread and returns `code`
int fail(int code, const char* fmt, ...);
// VSLAM_OK iff `device` is a usable sm_100 CUDA device (there is no CPU fallback)
int require_device(int device);

}  // namespace vslam

#define CUDA_TRY(expr)                                                                                       \
  do {                                                                                                       \
    const cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess)                                                                                   \
      return ::vslam::fail(-2 /* VSLAM_ERR_CUDA */, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                           __FILE__, __LINE__);                                                              \
  } while (0)
