// aligner_internal.h -- what the generator's C ABI reads of an aligner handle (same library, other translation unit):
// the per-correspondence results of its last run, on the device and -- once they have been brought back -- on the host.
#pragma once
#include <cstdint>

struct vslam_aligner;

namespace vslam {

struct AlignerResults {
  int n = 0;                        // correspondences of the last upload
  int device = 0;
  const double* d_errors = nullptr;     // device
  const uint8_t* d_inliers = nullptr;
  const double* h_errors = nullptr;     // pinned host copy (valid after fetch_aligner_results)
  const uint8_t* h_inliers = nullptr;
  double total_error = 0;           // of the last system read back
};

// makes the host copy current (one device -> host copy unless the fused converge already brought it back)
int fetch_aligner_results(vslam_aligner* h, AlignerResults* out);

}  // namespace vslam
