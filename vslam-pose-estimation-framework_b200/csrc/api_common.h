// api_common.h -- error plumbing shared by the C ABI translation units.
#pragma once

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: a no-op (one pointer test) unless a profiler injects itself

namespace vslam {

// records the message for vslam_last_error() on this thread and returns `code`
int fail(int code, const char* fmt, ...);
// VSLAM_OK iff `device` is a usable sm_100 CUDA device (there is no CPU fallback)
int require_device(int device);

// NVTX range over one C-ABI call, named like the reference's easy_profiler blocks where it has one
// (EASY_BLOCK("KeypointDetection" / "DescriptorExtraction" / "StereoMatching" ...), base_framepoint_generator.cpp:359,
// :433, stereo_framepoint_generator.cpp:137, pose_tracker_3d.cpp:91, :122, :186, :477): Nsight Systems shows the stages
// of a frame on the host timeline next to the kernels they launch
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

}  // namespace vslam

#define VSLAM_NVTX(name) ::vslam::NvtxRange _vslam_nvtx_range(name)

#define CUDA_TRY(expr)                                                                                       \
  do {                                                                                                       \
    const cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess)                                                                                   \
      return ::vslam::fail(-2 /* VSLAM_ERR_CUDA */, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                           __FILE__, __LINE__);                                                              \
  } while (0)
