// host_math.cpp -- host-side scalar pieces of the path that stay on the CPU by design (C++14, no CUDA):
//   * FAST threshold controller      (reference base_framepoint_generator.cpp:377-415, 440-459)
//   * detector region grid, bin grid (base_framepoint_generator.cpp:231-312)
//   * 6x6 complete-pivoting LU solve (Eigen::FullPivLU use at stereouv_aligner.cpp:199 / uvd_aligner.cpp:183)
//   * srrg_core::v2t                 (stereouv_aligner.cpp:200 / uvd_aligner.cpp:184; srrg_core is not vendored by
//                                     the reference -- restated from its published definition)
#include "host_math.h"

#include "gn_math.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <utility>

namespace vslam {

void detector_regions(int rows, int cols, int nv, int nh, HostRegion* out) {
  const double rows_per_detector = static_cast<double>(rows) / nv;
  const double cols_per_detector = static_cast<double>(cols) / nh;
  for (int r = 0; r < nv; ++r) {
    for (int c = 0; c < nh; ++c) {
      int grow_w = nh > 1 ? 2 : 0, grow_h = nv > 1 ? 2 : 0;   // overlap so no point is lost at region borders
      int shift_r = 0, shift_c = 0;
      if (r > 0) {
        shift_r = -grow_h;
        if (r < nv - 1) grow_h *= 2;   // interior regions overlap on both sides
      }
      if (c > 0) {
        shift_c = -grow_w;
        if (c < nh - 1) grow_w *= 2;
      }
      HostRegion& q = out[r * nh + c];   // cv::Rect(int, int, int, int) built from doubles: truncation
      q.x = static_cast<int>(std::round(c * cols_per_detector) + shift_c);
      q.y = static_cast<int>(std::round(r * rows_per_detector) + shift_r);
      q.w = static_cast<int>(cols_per_detector + grow_w);
      q.h = static_cast<int>(rows_per_detector + grow_h);
    }
  }
}

void bin_grid(int rows, int cols, int bin_size, int* rows_bin, int* cols_bin) {
  *cols_bin = static_cast<int>(std::floor(static_cast<double>(cols) / bin_size) + 1);
  *rows_bin = static_cast<int>(std::floor(static_cast<double>(rows) / bin_size) + 1);
}

double threshold_proposal(double threshold, int n_keypoints, double target, double tolerance, double maximum_change,
                          double threshold_minimum, double threshold_maximum) {
  const double delta = (static_cast<double>(n_keypoints) - target) / target;   // 100% loss -> -1, 100% gain -> +1
  if (delta < -tolerance) {
    const double change = std::max(delta, -maximum_change);
    threshold += std::min(change * threshold, -1.0);   // always lower by at least 1
    if (threshold < threshold_minimum) threshold = threshold_minimum;
  } else if (delta > tolerance) {
    const double change = std::min(delta, maximum_change);
    threshold += std::max(change * threshold, 1.0);    // always raise by at least 1
    if (threshold > threshold_maximum) threshold = threshold_maximum;
  }
  return threshold;
}

// Eigen::Quaternion<double>(Matrix3) as WorldMap::writeTrajectoryTUM uses it (world_map.cpp:235); Eigen is un-vendored:
// restated from Eigen 3.3's rotation-matrix assignment (largest of trace / diagonal element decides the branch)
void rotation_to_quaternion(const double R[9], double q[4]) {
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[7] - R[5]) * t;
    q[1] = (R[2] - R[6]) * t;
    q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (R[3 * k + j] - R[3 * j + k]) * t;
    q[j] = (R[3 * j + i] + R[3 * i + j]) * t;
    q[k] = (R[3 * k + i] + R[3 * i + k]) * t;
  }
}

// world_map.cpp:196-214: outfile << std::fixed << std::setprecision(9); twelve values row by row, each followed by " "
int format_trajectory_kitti(const double T[12], char* line, int capacity) {
  int n = 0;
  for (int i = 0; i < 12; ++i) n += std::snprintf(line + (n < capacity ? n : 0), n < capacity ? (size_t)(capacity - n) : 0, "%.9f ", T[i]);
  n += std::snprintf(line + (n < capacity ? n : 0), n < capacity ? (size_t)(capacity - n) : 0, "\n");
  return n;
}

// world_map.cpp:230-248: timestamp, translation x y z, orientation x y z w
int format_trajectory_tum(double timestamp_seconds, const double T[12], char* line, int capacity) {
  const double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
  double q[4];
  rotation_to_quaternion(R, q);
  return std::snprintf(line, capacity > 0 ? (size_t)capacity : 0, "%.9f %.9f %.9f %.9f %.9f %.9f %.9f %.9f \n",
                       timestamp_seconds, T[3], T[7], T[11], q[0], q[1], q[2], q[3]);
}

}  // namespace vslam
