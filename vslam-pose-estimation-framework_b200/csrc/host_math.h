// host_math.h -- host-side scalar pieces of the path (see host_math.cpp).
#pragma once

namespace vslam {

struct HostRegion {
  int x, y, w, h;
};

void detector_regions(int rows, int cols, int nv, int nh, HostRegion* out);
void bin_grid(int rows, int cols, int bin_size, int* rows_bin, int* cols_bin);
double threshold_proposal(double threshold, int n_keypoints, double target, double tolerance, double maximum_change,
                          double threshold_minimum, double threshold_maximum);
// solve6 / v2t / apply_update live in gn_math.h (shared with the device code)

}  // namespace vslam
