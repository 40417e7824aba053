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
void solve6(const double A[36], const double rhs[6], double x[6]);
void v2t(const double v[6], double T[12]);
// T <- v2t(dx) * T followed by the first-order re-orthonormalisation of the rotation block
void apply_update(const double dx[6], double T[12]);

}  // namespace vslam
