// Shape-only stand-in (tests/stubs/README.md).
#pragma once
#include <Eigen/Geometry>
#include <opencv2/opencv.hpp>
#include <string>
namespace srrg_core {
class PinholeImageMessage {
 public:
  const cv::Mat& image() const; const Eigen::Matrix<float, 3, 3>& cameraMatrix() const;
  const Eigen::Transform<float, 3, Eigen::Isometry>& offset() const; const Eigen::Transform<float, 3, Eigen::Isometry>& odometry() const;
  double timestamp() const; const std::string& topic() const; float depthScale() const;
};
}  // namespace srrg_core
