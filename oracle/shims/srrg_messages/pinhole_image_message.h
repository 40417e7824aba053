// oracle/shims -- stand-in for srrg_core::PinholeImageMessage (dataset playback; outside the hot path).
// TEST INFRASTRUCTURE ONLY.  A plain value holder: the reference's Camera(PinholeImageMessage*) constructor
// (src/types/camera.cpp:19) reads image(), cameraMatrix() and offset().
#pragma once
#include <Eigen/Geometry>
#include <opencv2/opencv.hpp>
#include <string>
namespace srrg_core {
class PinholeImageMessage {
 public:
  const cv::Mat& image() const { return _image; }
  void setImage(const cv::Mat& m) { _image = m; }
  const Eigen::Matrix<float, 3, 3>& cameraMatrix() const { return _camera_matrix; }
  void setCameraMatrix(const Eigen::Matrix<float, 3, 3>& k) { _camera_matrix = k; }
  const Eigen::Transform<float, 3, Eigen::Isometry>& offset() const { return _offset; }
  void setOffset(const Eigen::Transform<float, 3, Eigen::Isometry>& t) { _offset = t; }
  const Eigen::Transform<float, 3, Eigen::Isometry>& odometry() const { return _odometry; }
  double timestamp() const { return _timestamp; }
  const std::string& topic() const { return _topic; }
  float depthScale() const { return _depth_scale; }
  const char* className() const { return "PinholeImageMessage"; }
 private:
  cv::Mat _image;
  Eigen::Matrix<float, 3, 3> _camera_matrix;
  Eigen::Transform<float, 3, Eigen::Isometry> _offset, _odometry;
  double _timestamp = 0;
  std::string _topic;
  float _depth_scale = 1e-3f;
};
}  // namespace srrg_core
