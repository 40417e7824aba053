// Shape-only stand-in for srrg_core types (tests/stubs/README.md): only what the reference's headers name.
#pragma once
#include <Eigen/Geometry>
#include <opencv2/opencv.hpp>
namespace srrg_core {
typedef Eigen::Matrix<float, 6, 1> Vector6f; typedef Eigen::Matrix<double, 6, 1> Vector6d;
typedef Eigen::Matrix<float, 6, 6> Matrix6f; typedef Eigen::Matrix<double, 6, 6> Matrix6d;
typedef Eigen::Matrix<double, 3, 3> Matrix3d; typedef Eigen::Matrix<double, 3, 1> Vector3d;
template <class V> Eigen::Matrix<double, 3, 3> skew(const V&);
template <class V> Eigen::Transform<double, 3, Eigen::Isometry> v2t(const V&);
template <class T> Eigen::Matrix<double, 6, 1> t2v(const T&);
template <class M> cv::Mat toCv(const M&);
template <class T, int N> Eigen::Matrix<T, N, 1> fromCv(const cv::Vec<T, N>&);
}  // namespace srrg_core
