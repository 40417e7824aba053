// vslam_b200.hpp -- C++14 host layer above the C ABI (vslam_b200.h), mirroring the reference's plugin interface for
// the hot path with plain value types instead of its Eigen / OpenCV / srrg object graph, so that it builds with
// nothing but a C++14 compiler:
//
//   vslam::StereoFramePointGenerator::{configure (ctor), initialize, track, compute, recoverPoints}
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
  // track() / recoverPoints(): points() of a processed frame as the NEXT frame's track() reads them, and what the two
  // calls put into points() of the current frame
  std::vector<vslam_previous_point> previous_points;
  std::vector<vslam_track> tracks;
  std::vector<vslam_recovered_point> recovered;
  double average_descriptor_distance_tracking = 0;            // setAverageDescriptorDistanceTracking (on the previous frame)
  std::array<double, 12> world_to_camera_left{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}};
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
    if (!extract_features) {   // :126-133 only: both matchers are set up again, a track() attempt's pruning is forgotten
      check(vslam_fpg_reset_features(_handle), "StereoFramePointGenerator::initialize");
      _tracks_resident = false;
      return;
    }
    int32_t nl = 0, nr = 0;
    check(vslam_fpg_initialize(_handle, frame->intensity_image_left, frame->intensity_image_right, frame->image_step,
                               frame->status == Frame::Localizing, &nl, &nr), "StereoFramePointGenerator::initialize");
    fetch(0, nl, frame->keypoints_left, frame->descriptors_left);
    fetch(1, nr, frame->keypoints_right, frame->descriptors_right);
    _number_of_detected_keypoints = nl;
    _tracks_resident = false;
  }

  // base_framepoint_generator.h:164-165
  void setProjectionTrackingDistancePixels(const int32_t& v) { _projection_tracking_distance_pixels = v; }
  void setMaximumDescriptorDistanceTracking(const double& v) { _maximum_descriptor_distance_tracking = v; }

  // track(frame, frame_previous, camera_left_previous_in_current, lost_points, track_by_appearance):
  // frame->tracks = points() after the call; lost_points = positions in frame_previous->previous_points
  void track(Frame* frame, Frame* frame_previous, const std::array<double, 12>& camera_left_previous_in_current,
             std::vector<int32_t>& lost_points, const bool track_by_appearance = true) {
    if (!frame || !frame_previous) throw std::runtime_error("StereoFramePointGenerator::track|called with invalid frames");
    const std::vector<vslam_previous_point>& previous = frame_previous->previous_points;
    frame->tracks.resize(previous.size());
    lost_points.resize(previous.size());
    int32_t n = 0, n_lost = 0, n_landmarks = 0;
    check(vslam_fpg_track(_handle, previous.data(), (int32_t)previous.size(), camera_left_previous_in_current.data(),
                          track_by_appearance, _projection_tracking_distance_pixels, _maximum_descriptor_distance_tracking,
                          frame->tracks.data(), (int32_t)frame->tracks.size(), &n, lost_points.data(), &n_lost,
                          &n_landmarks, &frame_previous->average_descriptor_distance_tracking),
          "StereoFramePointGenerator::track");
    frame->tracks.resize(n);
    lost_points.resize(n_lost);
    _number_of_tracked_landmarks = n_landmarks;
    _tracks_resident = true;
  }

  void compute(Frame* frame) {
    if (!frame) throw std::runtime_error("StereoFramePointGenerator::compute|called with empty frame");
    // the tracks of this frame are still on the device unless the caller supplies its own list of points()
    const bool resident = _tracks_resident && frame->tracked_points.empty();
    frame->points.resize((size_t)_capacity + frame->tracked_points.size());
    int32_t n = 0, m = 0;
    check(vslam_fpg_compute(_handle, resident ? nullptr : frame->tracked_points.data(),
                            resident ? VSLAM_TRACKED_FROM_LAST_TRACK : (int32_t)frame->tracked_points.size(),
                            frame->points.data(), (int32_t)frame->points.size(), &n, &m),
          "StereoFramePointGenerator::compute");
    frame->points.resize(n);
    _number_of_new_points = m;
  }

  // recoverPoints(current_frame, lost_points): lost = the lost points (with landmark coordinates in `world`)
  void recoverPoints(Frame* current_frame, const std::vector<vslam_previous_point>& lost_points,
                     double minimum_depth_meters = 0.1, double maximum_depth_meters = 1000) const {
    if (!current_frame) throw std::runtime_error("StereoFramePointGenerator::recoverPoints|called with empty frame");
    current_frame->recovered.resize(lost_points.size());
    int32_t n = 0;
    check(vslam_fpg_recover_points(_handle, lost_points.data(), (int32_t)lost_points.size(),
                                   current_frame->world_to_camera_left.data(), minimum_depth_meters, maximum_depth_meters,
                                   _maximum_descriptor_distance_tracking, current_frame->recovered.data(),
                                   (int32_t)current_frame->recovered.size(), &n),
          "StereoFramePointGenerator::recoverPoints");
    current_frame->recovered.resize(n);
  }

  // PoseTracker3D::_prunePoints (pose_tracker_3d.cpp:437-472) after the aligner converged on frame->tracks: the rejected
  // tracks leave frame->tracks AND the device records that pre-load the bins of the next compute(), without an upload
  void pruneTracks(Frame* frame, vslam_aligner* aligner, double maximum_error_kernel) {
    if (!frame) throw std::runtime_error("StereoFramePointGenerator::pruneTracks|called with empty frame");
    std::vector<uint8_t> kept(frame->tracks.size() + 1);
    int32_t n = 0;
    check(vslam_fpg_prune_tracks(_handle, aligner, maximum_error_kernel, &n, kept.data()), "StereoFramePointGenerator::pruneTracks");
    size_t w = 0;
    for (size_t k = 0; k < frame->tracks.size(); ++k)
      if (kept[k]) frame->tracks[w++] = frame->tracks[k];
    frame->tracks.resize(w);
  }

  // PoseTracker3D::compute's per-frame order (pose_tracker_3d.cpp:80 initialize, :239 track, :355-357 StereoUVAligner
  // initialize + converge, :437-472 _prunePoints, :210 compute) as ONE device pass with points() of the previous frame
  // resident on the device (vslam_fpg_frame_step): frame->tracks, frame->points and frame->previous_points (= points() of
  // this frame, when parameters.publish_frame_points) are filled from the handle's pinned result block.
  vslam_frame_step_result trackFrame(Frame* frame, const std::array<double, 12>& previous_to_current_prior,
                                     const vslam_frame_step_parameters& parameters) {
    if (!frame) throw std::runtime_error("StereoFramePointGenerator::trackFrame|called with empty frame");
    vslam_frame_step_result r;
    check(vslam_fpg_frame_step(_handle, frame->intensity_image_left, frame->intensity_image_right, frame->image_step,
                               frame->status == Frame::Localizing, previous_to_current_prior.data(), &parameters, &r),
          "StereoFramePointGenerator::trackFrame");
    frame->tracks.assign(r.tracks, r.tracks + r.n_tracks);
    frame->points.assign(r.points, r.points + r.n_new_points);
    if (r.frame_points) frame->previous_points.assign(r.frame_points, r.frame_points + r.n_tracks + r.n_new_points);
    _number_of_detected_keypoints = r.n_left;
    _number_of_tracked_landmarks = r.n_tracked_landmarks;
    _number_of_new_points = r.n_matches;
    _tracks_resident = false;
    return r;
  }
  // the images of the NEXT frame, uploaded while the current one runs (vslam_fpg_frame_step_prefetch); the trackFrame()
  // that consumes them is called with frame->intensity_image_left == frame->intensity_image_right == nullptr
  void prefetchFrame(const uint8_t* left, const uint8_t* right, size_t image_step) {
    check(vslam_fpg_frame_step_prefetch(_handle, left, right, image_step), "StereoFramePointGenerator::prefetchFrame");
  }
  // landmark estimates of the points() the device holds (stereouv_aligner.cpp:43-51), one entry per point of the frame
  // that just returned; read by the next trackFrame()
  void setLandmarkEstimates(const std::vector<vslam_landmark_estimate>& estimates) {
    check(vslam_fpg_frame_step_set_landmark_estimates(_handle, estimates.data(), (int32_t)estimates.size()),
          "StereoFramePointGenerator::setLandmarkEstimates");
  }
  // a new sequence: the next trackFrame() has no previous points
  void resetSequence() { check(vslam_fpg_frame_step_reset(_handle), "StereoFramePointGenerator::resetSequence"); }

  int numberOfTrackedLandmarks() const { return _number_of_tracked_landmarks; }

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
  int32_t _projection_tracking_distance_pixels = 0;     // base_framepoint_generator.h:221
  double _maximum_descriptor_distance_tracking = 0;     // :224
  int _number_of_tracked_landmarks = 0;                 // :227
  bool _tracks_resident = false;
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
  vslam_aligner* handle() { return _handle; }
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

// The landmark loop of PoseTracker3D::_updatePoints (reference src/position_tracking/pose_tracker_3d.cpp:475-549) as ONE
// call: Landmark::update (src/types/landmark.cpp:66-152) for every landmark of the frame, one warp each.
class LandmarkOptimizer {
 public:
  LandmarkOptimizer(int32_t max_landmarks, int32_t max_measurements, int32_t max_frames, int device = 0) {
    check(vslam_landmark_optimizer_create(max_landmarks, max_measurements, max_frames, device, &_handle),
          "LandmarkOptimizer::LandmarkOptimizer");
  }
  ~LandmarkOptimizer() { vslam_landmark_optimizer_destroy(_handle); }
  LandmarkOptimizer(const LandmarkOptimizer&) = delete;
  LandmarkOptimizer& operator=(const LandmarkOptimizer&) = delete;

  // LandmarkParameters (src/types/parameters.h:111-114)
  uint32_t maximum_number_of_iterations = 100;
  double maximum_error_squared_meters = 5 * 5;

  // offsets: CSR over `measurements` (size n + 1); poses: [n_frames][12] row-major 3x4; world / number_of_updates in place
  void update(const std::vector<int32_t>& offsets, const std::vector<vslam_landmark_measurement>& measurements,
              const std::vector<double>& world_to_camera_left, const std::vector<double>& camera_left_to_world,
              std::vector<double>& world_coordinates, std::vector<uint32_t>& number_of_updates,
              std::vector<uint8_t>* outcome = nullptr) {
    const int32_t n = static_cast<int32_t>(offsets.size()) - 1;
    if (n < 0 || world_coordinates.size() != 3 * static_cast<size_t>(n) || number_of_updates.size() != static_cast<size_t>(n) ||
        world_to_camera_left.size() != camera_left_to_world.size() || world_to_camera_left.size() % 12 != 0)
      throw std::runtime_error("LandmarkOptimizer::update|inconsistent buffer sizes");
    if (outcome) outcome->resize(n);
    check(vslam_landmark_optimizer_update(_handle, n, offsets.data(), measurements.data(),
                                          static_cast<int32_t>(world_to_camera_left.size() / 12), world_to_camera_left.data(),
                                          camera_left_to_world.data(), maximum_number_of_iterations,
                                          maximum_error_squared_meters, world_coordinates.data(), number_of_updates.data(),
                                          outcome ? outcome->data() : nullptr, nullptr),
          "LandmarkOptimizer::update");
  }

 private:
  vslam_landmark_optimizer* _handle = nullptr;
};

// WorldMap::writeTrajectoryKITTI / writeTrajectoryTUM (reference src/types/world_map.cpp:183-252); poses: [n][12] row-major
inline void writeTrajectoryKITTI(const std::string& filename, const std::vector<double>& robot_to_world) {
  if (vslam_write_trajectory(filename.c_str(), VSLAM_TRAJECTORY_KITTI, static_cast<int32_t>(robot_to_world.size() / 12),
                             robot_to_world.data(), nullptr) != VSLAM_OK)
    throw std::runtime_error(std::string("writeTrajectoryKITTI|") + vslam_last_error());
}
inline void writeTrajectoryTUM(const std::string& filename, const std::vector<double>& timestamps_seconds,
                               const std::vector<double>& robot_to_world) {
  if (timestamps_seconds.size() * 12 != robot_to_world.size())
    throw std::runtime_error("writeTrajectoryTUM|one timestamp per pose expected");
  if (vslam_write_trajectory(filename.c_str(), VSLAM_TRAJECTORY_TUM, static_cast<int32_t>(timestamps_seconds.size()),
                             robot_to_world.data(), timestamps_seconds.data()) != VSLAM_OK)
    throw std::runtime_error(std::string("writeTrajectoryTUM|") + vslam_last_error());
}

}  // namespace vslam
