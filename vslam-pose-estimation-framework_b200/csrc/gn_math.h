// gn_math.h -- the 6x6 Gauss-Newton step shared by the host driver (host_math.cpp, g++ -ffp-contract=off) and the
// fused device solver (aligner.cu, nvcc -fmad=false): identical expressions, no contraction on either side, so the
// pose sequences of vslam_aligner_converge and vslam_aligner_converge_fused are bit-identical.
//   solve6        Eigen::FullPivLU<Matrix6>::solve as used at reference stereouv_aligner.cpp:199 / uvd_aligner.cpp:183
//   v2t           srrg_core::v2t (stereouv_aligner.cpp:200); srrg_core is not vendored by the reference
//   apply_update  T <- v2t(dx) * T, then R -= 0.5 * R * (R^T R - I)   (stereouv_aligner.cpp:200-206)
#pragma once

#include <math.h>

#ifdef __CUDACC__
#define VSLAM_HD __host__ __device__
#else
#define VSLAM_HD
#endif

namespace vslam {

template <typename T>
VSLAM_HD inline void swap_values(T& a, T& b) {
  const T t = a;
  a = b;
  b = t;
}

VSLAM_HD inline void solve6(const double A_in[36], const double rhs[6], double x[6]) {
  double A[36], b[6];
  int perm[6];
  for (int i = 0; i < 36; ++i) A[i] = A_in[i];
  for (int i = 0; i < 6; ++i) b[i] = rhs[i];
  for (int i = 0; i < 6; ++i) perm[i] = i;
  int rank = 6;
  double maxpivot = 0;
  for (int k = 0; k < 6; ++k) {
    int pr = k, pc = k;
    double biggest = -1;
    for (int i = k; i < 6; ++i)
      for (int j = k; j < 6; ++j)
        if (fabs(A[i * 6 + j]) > biggest) {
          biggest = fabs(A[i * 6 + j]);
          pr = i;
          pc = j;
        }
    if (biggest == 0) {
      rank = k;
      break;
    }
    if (biggest > maxpivot) maxpivot = biggest;
    if (biggest > maxpivot) maxpivot = biggest;
    if (pr != k) {
      for (int j = 0; j < 6; ++j) swap_values(A[k * 6 + j], A[pr * 6 + j]);
      swap_values(b[k], b[pr]);
    }
    if (pc != k) {
      for (int i = 0; i < 6; ++i) swap_values(A[i * 6 + k], A[i * 6 + pc]);
      swap_values(perm[k], perm[pc]);
    }
    for (int i = k + 1; i < 6; ++i) {
      const double f = A[i * 6 + k] / A[k * 6 + k];
      A[i * 6 + k] = f;
      for (int j = k + 1; j < 6; ++j) A[i * 6 + j] -= f * A[k * 6 + j];
      b[i] -= f * b[k];
    }
  }
  {  /* Eigen::FullPivLU::rank(): only pivots above |largest pivot| * epsilon * size are used by solve() */
    int r = 0;
    for (int i = 0; i < rank; ++i) r += fabs(A[i * 6 + i]) > maxpivot * (2.220446049250313e-16 * 6);
    rank = r;
  }
  double y[6] = {0, 0, 0, 0, 0, 0};
  for (int i = rank - 1; i >= 0; --i) {
    double s = b[i];
    for (int j = i + 1; j < rank; ++j) s -= A[i * 6 + j] * y[j];
    y[i] = s / A[i * 6 + i];
  }
  for (int i = 0; i < 6; ++i) x[perm[i]] = y[i];
}

// Eigen::FullPivLU<Matrix3>::solve as used by Landmark::update (reference src/types/landmark.cpp:136): the same
// complete-pivoting elimination as solve6, on 3 x 3
VSLAM_HD inline void solve3(const double A_in[9], const double rhs[3], double x[3]) {
  double A[9], b[3];
  int perm[3] = {0, 1, 2};
  for (int i = 0; i < 9; ++i) A[i] = A_in[i];
  for (int i = 0; i < 3; ++i) b[i] = rhs[i];
  int rank = 3;
  double maxpivot = 0;
  for (int k = 0; k < 3; ++k) {
    int pr = k, pc = k;
    double biggest = -1;
    for (int i = k; i < 3; ++i)
      for (int j = k; j < 3; ++j)
        if (fabs(A[i * 3 + j]) > biggest) {
          biggest = fabs(A[i * 3 + j]);
          pr = i;
          pc = j;
        }
    if (biggest == 0) {
      rank = k;
      break;
    }
    if (biggest > maxpivot) maxpivot = biggest;
    if (biggest > maxpivot) maxpivot = biggest;
    if (pr != k) {
      for (int j = 0; j < 3; ++j) swap_values(A[k * 3 + j], A[pr * 3 + j]);
      swap_values(b[k], b[pr]);
    }
    if (pc != k) {
      for (int i = 0; i < 3; ++i) swap_values(A[i * 3 + k], A[i * 3 + pc]);
      swap_values(perm[k], perm[pc]);
    }
    for (int i = k + 1; i < 3; ++i) {
      const double f = A[i * 3 + k] / A[k * 3 + k];
      A[i * 3 + k] = f;
      for (int j = k + 1; j < 3; ++j) A[i * 3 + j] -= f * A[k * 3 + j];
      b[i] -= f * b[k];
    }
  }
  {  /* Eigen::FullPivLU::rank(): only pivots above |largest pivot| * epsilon * size are used by solve() */
    int r = 0;
    for (int i = 0; i < rank; ++i) r += fabs(A[i * 3 + i]) > maxpivot * (2.220446049250313e-16 * 3);
    rank = r;
  }
  double y[3] = {0, 0, 0};
  for (int i = rank - 1; i >= 0; --i) {
    double s = b[i];
    for (int j = i + 1; j < rank; ++j) s -= A[i * 3 + j] * y[j];
    y[i] = s / A[i * 3 + i];
  }
  for (int i = 0; i < 3; ++i) x[perm[i]] = y[i];
}

VSLAM_HD inline void v2t(const double v[6], double T[12]) {
  double qx = v[3], qy = v[4], qz = v[5], qw;
  const double n2 = qx * qx + qy * qy + qz * qz;
  if (n2 < 1) {
    qw = sqrt(1 - n2);
  } else {
    const double n = sqrt(n2);
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

VSLAM_HD inline void apply_update(const double dx[6], double T[12]) {
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
  for (int i = 0; i < 12; ++i) T[i] = Tn[i];
  for (int i = 0; i < 3; ++i)    // R -= 0.5 * R * (R^T R - I)
    for (int j = 0; j < 3; ++j)
      T[4 * i + j] = Tn[4 * i + j] - 0.5 * (Tn[4 * i] * E[j] + Tn[4 * i + 1] * E[3 + j] + Tn[4 * i + 2] * E[6 + j]);
}


}  // namespace vslam
