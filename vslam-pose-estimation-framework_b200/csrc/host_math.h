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
// one line of WorldMap::writeTrajectoryKITTI / writeTrajectoryTUM (reference src/types/world_map.cpp:183-252): std::fixed,
// setprecision(9), every value followed by a blank; returns the characters written (snprintf semantics)
int format_trajectory_kitti(const double robot_to_world[12], char* line, int capacity);
int format_trajectory_tum(double timestamp_seconds, const double robot_to_world[12], char* line, int capacity);
// Eigen::Quaternion(Matrix3) (world_map.cpp:235): q = (x, y, z, w)
void rotation_to_quaternion(const double R[9], double q[4]);
// solve6 / solve3 / v2t / apply_update live in gn_math.h (shared with the device code)

}  // namespace vslam
