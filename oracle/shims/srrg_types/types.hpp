// oracle/shims -- functional stand-in for srrg_core's types.hpp (branch marchless, un-vendored).
// TEST INFRASTRUCTURE ONLY (see oracle/shims/Eigen/Core).  Restated from srrg_core's published definitions:
//   skew(p)  = [[0, -z, y], [z, 0, -x], [-y, x, 0]]
//   v2t(v)   : translation v[0:3]; rotation = quaternion (w, v[3:6]) with w = sqrt(1 - |q|^2) if |q|^2 < 1,
//              else (0, q / |q|); t.linear() = q.toRotationMatrix()
//   t2v(t)   : inverse of v2t (translation; vector part of the normalised quaternion, sign chosen so that w >= 0)
// The -2 * skew(p) rotation Jacobian of stereouv_aligner.cpp:149 is the derivative of exactly this v2t.
#pragma once
#include <Eigen/Geometry>
#include <opencv2/opencv.hpp>
namespace srrg_core {
typedef Eigen::Matrix<float, 6, 1> Vector6f; typedef Eigen::Matrix<double, 6, 1> Vector6d;
typedef Eigen::Matrix<float, 6, 6> Matrix6f; typedef Eigen::Matrix<double, 6, 6> Matrix6d;
typedef Eigen::Matrix<double, 3, 3> Matrix3d; typedef Eigen::Matrix<double, 3, 1> Vector3d;

template <class D, class S>
Eigen::Matrix<S, 3, 3> skew(const Eigen::MatrixBase<D, S, 3, 1>& p) {
  Eigen::Matrix<S, 3, 3> s;
  s << S(0), -p(2), p(1),
       p(2), S(0), -p(0),
       -p(1), p(0), S(0);
  return s;
}

template <class D, class S>
Eigen::Transform<S, 3, Eigen::Isometry> v2t(const Eigen::MatrixBase<D, S, 6, 1>& v) {
  Eigen::Transform<S, 3, Eigen::Isometry> t;
  t.translation() = v.template head<3>();
  S w = v(3) * v(3) + v(4) * v(4) + v(5) * v(5);      // = v.block<3,1>(3,0).squaredNorm()
  if (w < S(1)) {
    w = std::sqrt(S(1) - w);
    t.linear() = Eigen::Quaternion<S>(w, v(3), v(4), v(5)).toRotationMatrix();
  } else {
    Eigen::Matrix<S, 3, 1> q(v(3), v(4), v(5));
    q.normalize();
    t.linear() = Eigen::Quaternion<S>(S(0), q(0), q(1), q(2)).toRotationMatrix();
  }
  return t;
}

template <class S, int Mode, int O>
Eigen::Matrix<S, 6, 1> t2v(const Eigen::Transform<S, 3, Mode, O>& t) {
  Eigen::Matrix<S, 6, 1> v;
  const Eigen::Matrix<S, 3, 1> tr = t.translation();
  Eigen::Quaternion<S> q(t.linear());
  q.normalize();
  const S sign = q.w() < S(0) ? S(-1) : S(1);
  v(0) = tr(0); v(1) = tr(1); v(2) = tr(2);
  v(3) = sign * q.x(); v(4) = sign * q.y(); v(5) = sign * q.z();
  return v;
}

template <class D, class S, int R, int C>
cv::Mat toCv(const Eigen::MatrixBase<D, S, R, C>& m) {
  cv::Mat r(R, C, CV_64FC1);
  for (int i = 0; i < R; ++i)
    for (int j = 0; j < C; ++j) r.at<double>(i, j) = static_cast<double>(m(i, j));
  return r;
}
template <class T, int N>
Eigen::Matrix<T, N, 1> fromCv(const cv::Vec<T, N>& v) {
  Eigen::Matrix<T, N, 1> r;
  for (int i = 0; i < N; ++i) r(i) = v[i];
  return r;
}
}  // namespace srrg_core
