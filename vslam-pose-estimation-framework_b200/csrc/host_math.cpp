// host_math.cpp -- host-side scalar pieces of the path that stay on the CPU by design (C++14, no CUDA):
//   * FAST threshold controller      (reference base_framepoint_generator.cpp:377-415, 440-459)
//   * detector region grid, bin grid (base_framepoint_generator.cpp:231-312)
//   * 6x6 complete-pivoting LU solve (Eigen::FullPivLU use at stereouv_aligner.cpp:199 / uvd_aligner.cpp:183)
//   * srrg_core::v2t                 (stereouv_aligner.cpp:200 / uvd_aligner.cpp:184; srrg_core is not vendored by
//                                     the reference -- restated from its published definition)
#include "host_math.h"

#include <algorithm>
#include <cmath>
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

void solve6(const double A_in[36], const double rhs[6], double x[6]) {
  double A[36], b[6];
  int perm[6];
  std::memcpy(A, A_in, sizeof(A));
  std::memcpy(b, rhs, sizeof(b));
  for (int i = 0; i < 6; ++i) perm[i] = i;
  int rank = 6;
  for (int k = 0; k < 6; ++k) {
    int pr = k, pc = k;
    double biggest = -1;
    for (int i = k; i < 6; ++i)
      for (int j = k; j < 6; ++j)
        if (std::fabs(A[i * 6 + j]) > biggest) {
          biggest = std::fabs(A[i * 6 + j]);
          pr = i;
          pc = j;
        }
    if (biggest == 0) {
      rank = k;
      break;
    }
    if (pr != k) {
      for (int j = 0; j < 6; ++j) std::swap(A[k * 6 + j], A[pr * 6 + j]);
      std::swap(b[k], b[pr]);
    }
    if (pc != k) {
      for (int i = 0; i < 6; ++i) std::swap(A[i * 6 + k], A[i * 6 + pc]);
      std::swap(perm[k], perm[pc]);
    }
    for (int i = k + 1; i < 6; ++i) {
      const double f = A[i * 6 + k] / A[k * 6 + k];
      A[i * 6 + k] = f;
      for (int j = k + 1; j < 6; ++j) A[i * 6 + j] -= f * A[k * 6 + j];
      b[i] -= f * b[k];
    }
  }
  double y[6] = {0, 0, 0, 0, 0, 0};
  for (int i = rank - 1; i >= 0; --i) {
    double s = b[i];
    for (int j = i + 1; j < rank; ++j) s -= A[i * 6 + j] * y[j];
    y[i] = s / A[i * 6 + i];
  }
  for (int i = 0; i < 6; ++i) x[perm[i]] = y[i];
}

void v2t(const double v[6], double T[12]) {
  double qx = v[3], qy = v[4], qz = v[5], qw;
  const double n2 = qx * qx + qy * qy + qz * qz;
  if (n2 < 1) {
    qw = std::sqrt(1 - n2);
  } else {
    const double n = std::sqrt(n2);
    qx /= n;
    qy /= n;
    qz /= n;
    qw = 0;
  }
  const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;   // Eigen::Quaternion::toRotationMatrix
  const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
  const double txx = tx * qx, txy = ty * qx, txz = tz * qx;
  const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  T[0] = 1 - (tyy + tzz); T[1] = txy - twz;       T[2] = txz + twy;        T[3] = v[0];
  T[4] = txy + twz;       T[5] = 1 - (txx + tzz); T[6] = tyz - twx;        T[7] = v[1];
  T[8] = txz - twy;       T[9] = tyz + twx;       T[10] = 1 - (txx + tyy); T[11] = v[2];
}

void apply_update(const double dx[6], double T[12]) {
  double D[12], Tn[12];
  v2t(dx, D);
  for (int i = 0; i < 3; ++i) {   // Tn = D * T
    for (int j = 0; j < 3; ++j) Tn[4 * i + j] = D[4 * i] * T[j] + D[4 * i + 1] * T[4 + j] + D[4 * i + 2] * T[8 + j];
    Tn[4 * i + 3] = D[4 * i] * T[3] + D[4 * i + 1] * T[7] + D[4 * i + 2] * T[11] + D[4 * i + 3];
  }
  double E[9];   // R^T R - I
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      E[3 * i + j] = Tn[i] * Tn[j] + Tn[4 + i] * Tn[4 + j] + Tn[8 + i] * Tn[8 + j];
      if (i == j) E[3 * i + j] -= 1;
    }
  std::memcpy(T, Tn, sizeof(double) * 12);
  for (int i = 0; i < 3; ++i)    // R -= 0.5 * R * (R^T R - I)
    for (int j = 0; j < 3; ++j)
      T[4 * i + j] = Tn[4 * i + j] - 0.5 * (Tn[4 * i] * E[j] + Tn[4 * i + 1] * E[3 + j] + Tn[4 * i + 2] * E[6 + j]);
}

}  // namespace vslam
