// vslam_b200.hpp -- C++14 host layer above the C ABI (vslam_b200.h), mirroring the reference's plugin interface for
// the hot path with plain value types instead of its Eigen / OpenCV / srrg object graph, so that it builds with
// nothing but a C++14 compiler:
//
//   vslam::StereoFramePointGenerator::{configure (ctor), initialize(Frame&), compute(Frame&)}
//        <- proslam::BaseFramePointGenerator / StereoFramePointGenerator
//           (reference src/framepoint_generation/base_framepoint_generator.h:119-151)
//   vslam::StereoUVAligner / vslam::UVDAligner::{initialize, linearize, oneRound, converge, errors, inliers, ...}
//        <- proslam::BaseAligner / BaseFrameAligner (reference src/aligners/base_aligner.h:26-48,
//           base_frame_aligner.h:20-27)
//
// Same method names, argument meaning and error behaviour (std::runtime_error, as the reference throws and
// executables/app.cpp:128 catches).  The adapters in adapters/ are the variant that derives from the reference's
// own classes.
#pragma once

#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "vslam_b200.h"

namespace vslam {

inline void check(int status, const char* who) {
  if (status != VSLAM_OK) throw std::runtime_error(std::string(who) + "|" + vslam_last_error());
}

// proslam::Frame reduced to what the hot path reads and writes (src/types/frame.h:64-67,88-89,114-122)
struct Frame {
  enum Status { Localizing, Tracking };
  Status status = Localizing;
  const uint8_t* intensity_image_left = nullptr;
  const uint8_t* intensity_image_right = nullptr;
  size_t image_step = 0;                                   // bytes between rows
  std::vector<vslam_keypoint> keypoints_left, keypoints_right;
  std::vector<uint8_t> descriptors_left, descriptors_right;   // n x 32
  std::vector<vslam_tracked_point> tracked_points;            // points() before compute()
  std::vector<vslam_framepoint> points;                       // what compute() appends to points()
};

class StereoFramePointGenerator {
 public:
  StereoFramePointGenerator(const vslam_fpg_config& parameters, int device = 0) {
    check(vslam_fpg_create(&parameters, device, &_handle), "StereoFramePointGenerator::configure");
    int32_t n = 0;
    check(vslam_fpg_info(_handle, &n, nullptr, &_rows_bin, &_cols_bin, &_target_number_of_keypoints), "configure");
    _number_of_detectors = n;
    _capacity = parameters.max_keypoints_per_image > 0 ? parameters.max_keypoints_per_image
                                                       : (4 * _target_number_of_keypoints > 4096 ? 4 * _target_number_of_keypoints : 4096);
    if (_capacity > 65535) _capacity = 65535;
  }
  ~StereoFramePointGenerator() { vslam_fpg_destroy(_handle); }
  StereoFramePointGenerator(const StereoFramePointGenerator&) = delete;
  StereoFramePointGenerator& operator=(const StereoFramePointGenerator&) = delete;

  void initialize(Frame* frame, const bool& extract_features = true) {
    if (!frame) throw std::runtime_error("StereoFramePointGenerator::initialize|called with empty frame");
    if (!extract_features) return;
    int32_t nl = 0, nr = 0;
    check(vslam_fpg_initialize(_handle, frame->intensity_image_left, frame->intensity_image_right, frame->image_step,
                               frame->status == Frame::Localizing, &nl, &nr), "StereoFramePointGenerator::initialize");
    fetch(0, nl, frame->keypoints_left, frame->descriptors_left);
    fetch(1, nr, frame->keypoints_right, frame->descriptors_right);
    _number_of_detected_keypoints = nl;
  }

  void compute(Frame* frame) {
    if (!frame) throw std::runtime_error("StereoFramePointGenerator::compute|called with empty frame");
    frame->points.resize((size_t)_capacity + frame->tracked_points.size());
    int32_t n = 0, m = 0;
    check(vslam_fpg_compute(_handle, frame->tracked_points.data(), (int32_t)frame->tracked_points.size(),
                            frame->points.data(), (int32_t)frame->points.size(), &n, &m),
          "StereoFramePointGenerator::compute");
    frame->points.resize(n);
    _number_of_new_points = m;
  }

  int targetNumberOfKeypoints() const { return _target_number_of_keypoints; }
  int numberOfDetectedKeypoints() const { return _number_of_detected_keypoints; }
  int numberOfNewPoints() const { return _number_of_new_points; }
  double meanDetectorThreshold() const {
    std::vector<double> t(VSLAM_MAX_DETECTOR_REGIONS);
    vslam_fpg_get_thresholds(_handle, t.data());
    double s = 0;
    for (int i = 0; i < _number_of_detectors; ++i) s += t[i];
    return s / _number_of_detectors;
  }
  vslam_fpg* handle() { return _handle; }

 private:
  void fetch(int side, int32_t n, std::vector<vslam_keypoint>& k, std::vector<uint8_t>& d) {
    k.resize(n);
    d.resize((size_t)n * VSLAM_DESCRIPTOR_BYTES);
    int32_t got = 0;
    check(vslam_fpg_get_features(_handle, side, k.data(), d.data(), n, &got), "StereoFramePointGenerator::initialize");
  }
  vslam_fpg* _handle = nullptr;
  int32_t _rows_bin = 0, _cols_bin = 0, _target_number_of_keypoints = 0;
  int _number_of_detectors = 1, _capacity = 0, _number_of_detected_keypoints = 0, _number_of_new_points = 0;
};

// BaseFrameAligner over caller-provided correspondence buffers (what ::initialize leaves behind)
template <int Kind>
class FrameAligner {
 public:
  explicit FrameAligner(const vslam_aligner_parameters& parameters, int32_t max_points = 1 << 16, int device = 0)
      : _parameters(parameters) {
    check(vslam_aligner_create(Kind, max_points, device, &_handle), "Aligner::Aligner");
    _previous_to_current = {{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}};
  }
  ~FrameAligner() { vslam_aligner_destroy(_handle); }
  FrameAligner(const FrameAligner&) = delete;
  FrameAligner& operator=(const FrameAligner&) = delete;

  void initialize(int32_t n, const double* moving, const double* fixed, const double* omega,
                  const double* weights_translation, const double K[9], const double baseline[3], int32_t rows,
                  int32_t cols, double minimum_reliable_depth_meters, const std::array<double, 12>& previous_to_current) {
    check(vslam_aligner_upload(_handle, n, moving, fixed, omega, weights_translation, K, baseline, rows, cols,
                               minimum_reliable_depth_meters), "Aligner::initialize");
    _number_of_measurements = n;
    _previous_to_current = previous_to_current;
  }
  void linearize(const bool& ignore_outliers) {
    check(vslam_aligner_linearize(_handle, _previous_to_current.data(), ignore_outliers, _parameters.maximum_error_kernel,
                                  &_system), "Aligner::linearize");
  }
  void oneRound(const bool& ignore_outliers) {
    check(vslam_aligner_one_round(_handle, &_parameters, ignore_outliers, _previous_to_current.data(), &_system),
          "Aligner::oneRound");
  }
  void converge() {
    int32_t ok = 0, rounds = 0;
    check(vslam_aligner_converge_fused(_handle, &_parameters, _previous_to_current.data(), &_system, _information_matrix.data(),
                                 &ok, &rounds), "Aligner::converge");
    _has_system_converged = ok != 0;
    _number_of_rounds = rounds;
  }
  std::vector<double> errors() const {
    std::vector<double> e(_number_of_measurements);
    check(vslam_aligner_download(_handle, e.data(), nullptr), "Aligner::errors");
    return e;
  }
  std::vector<bool> inliers() const {
    std::vector<uint8_t> b(_number_of_measurements);
    check(vslam_aligner_download(_handle, nullptr, b.data()), "Aligner::inliers");
    return std::vector<bool>(b.begin(), b.end());
  }
  int numberOfInliers() const { return _system.number_of_inliers; }
  int numberOfOutliers() const { return _system.number_of_outliers; }
  int numberOfCorrespondences() const { return _number_of_measurements; }
  double totalError() const { return _system.total_error; }
  double averageError() const { return _system.total_error / _number_of_measurements; }
  bool hasSystemConverged() const { return _has_system_converged; }
  int numberOfRounds() const { return _number_of_rounds; }
  const std::array<double, 12>& previousToCurrent() const { return _previous_to_current; }
  const vslam_linear_system& system() const { return _system; }
  vslam_aligner_parameters* parameters() { return &_parameters; }

 private:
  vslam_aligner* _handle = nullptr;
  vslam_aligner_parameters _parameters;
  vslam_linear_system _system = {};
  std::array<double, 12> _previous_to_current;
  std::array<double, 36> _information_matrix = {};
  int32_t _number_of_measurements = 0;
  bool _has_system_converged = false;
  int _number_of_rounds = 0;
};

typedef FrameAligner<VSLAM_ALIGNER_STEREO_UV> StereoUVAligner;
typedef FrameAligner<VSLAM_ALIGNER_UVD> UVDAligner;

}  // namespace vslam
