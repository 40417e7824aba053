// gpu_stereo_framepoint_generator.cpp -- see the header.  Every numbered comment cites the reference line whose
// effect the call reproduces (reference src/framepoint_generation/stereo_framepoint_generator.cpp unless noted).
#include "gpu_stereo_framepoint_generator.h"

#include "types/landmark.h"

#include <cstring>
#include <stdexcept>

namespace proslam {

GpuStereoFramePointGenerator::GpuStereoFramePointGenerator(StereoFramePointGeneratorParameters* parameters_,
                                                           int cuda_device_)
    : StereoFramePointGenerator(parameters_), _stereo_parameters(parameters_), _cuda_device(cuda_device_) {}

GpuStereoFramePointGenerator::~GpuStereoFramePointGenerator() { vslam_fpg_destroy(_handle); }

void GpuStereoFramePointGenerator::check(int status_) const {
  if (status_ != VSLAM_OK) throw std::runtime_error(std::string("GpuStereoFramePointGenerator|") + vslam_last_error());
}

void GpuStereoFramePointGenerator::configure() {
  // host-side bookkeeping of the base classes (bin grid, feature matchers used by the inherited track()): :16-60
  StereoFramePointGenerator::configure();

  const StereoFramePointGeneratorParameters* p = _stereo_parameters;
  vslam_fpg_config c = {};
  c.rows = _number_of_rows_image;
  c.cols = _number_of_cols_image;
  c.target_number_of_keypoints_tolerance = p->target_number_of_keypoints_tolerance;
  c.detector_threshold_minimum = p->detector_threshold_minimum;
  c.detector_threshold_maximum = p->detector_threshold_maximum;
  c.detector_threshold_maximum_change = p->detector_threshold_maximum_change;
  c.number_of_detectors_vertical = p->number_of_detectors_vertical;
  c.number_of_detectors_horizontal = p->number_of_detectors_horizontal;
  c.enable_keypoint_binning = p->enable_keypoint_binning;
  c.bin_size_pixels = p->bin_size_pixels;
  c.maximum_matching_distance_triangulation = p->maximum_matching_distance_triangulation;
  c.minimum_disparity_pixels = p->minimum_disparity_pixels;
  c.maximum_epipolar_search_offset_pixels = p->maximum_epipolar_search_offset_pixels;
  c.fx = _f_x; c.fy = _f_y; c.cx = _c_x; c.cy = _c_y; c.bx = _b_x;           // :26-38
  // after the base configure() the parameter holds what the reference really instantiated (base :184-224): "ORB" for
  // ORB / ORB-256 / BRIEF-256 and for BRIEF without opencv_contrib; "BRIEF" = xfeatures2d::BriefDescriptorExtractor(32)
  // when SRRG_PROSLAM_HAS_OPENCV_CONTRIB is defined.  parameters.cpp:341 never parses detector_type in stereo mode: FAST.
  if (p->detector_type != "FAST")
    throw std::runtime_error("GpuStereoFramePointGenerator::configure|only the FAST detector runs on the GPU");
  if (p->descriptor_type == "BRIEF") {
    if (!_brief_tests_set)
      throw std::runtime_error("GpuStereoFramePointGenerator::configure|descriptor_type BRIEF needs setBriefTests() "
                               "(the 256 x 4 table of opencv_contrib's generated_32.i, tools/parse_brief_generated.py)");
    c.descriptor_type = VSLAM_DESCRIPTOR_BRIEF;
    c.brief_tests = _brief_tests;
  } else if (p->descriptor_type != "ORB") {
    throw std::runtime_error("GpuStereoFramePointGenerator::configure|descriptor_type " + p->descriptor_type +
                             " does not run on the GPU (ORB-256 and BRIEF-32 do)");
  }
  vslam_fpg_destroy(_handle);
  _handle = nullptr;
  check(vslam_fpg_create(&c, _cuda_device, &_handle));
  check(vslam_fpg_set_profiling(_handle, 1));
}

void GpuStereoFramePointGenerator::initialize(Frame* frame_, const bool& extract_features_) {
  if (!frame_) throw std::runtime_error("StereoFramePointGenerator::initialize|called with empty frame");   // :75-78
  if (extract_features_) {
    const cv::Mat& left = frame_->intensityImageLeft();
    const cv::Mat& right = frame_->intensityImageRight();
    if (left.type() != CV_8UC1 || right.type() != CV_8UC1 || left.step != right.step)
      throw std::runtime_error("GpuStereoFramePointGenerator::initialize|expected two CV_8UC1 images of equal step");
    int32_t n_left = 0, n_right = 0;
    // :85-125 detectKeypoints x2, adjustDetectorThresholds, computeDescriptors x2, triangulation distance
    check(vslam_fpg_initialize(_handle, left.data, right.data, left.step, frame_->status() == Frame::Localizing,
                               &n_left, &n_right));
    for (int side = 0; side < 2; ++side) {
      std::vector<cv::KeyPoint>& keypoints = side == 0 ? frame_->keypointsLeft() : frame_->keypointsRight();
      cv::Mat& descriptors = side == 0 ? frame_->descriptorsLeft() : frame_->descriptorsRight();
      const int32_t n = side == 0 ? n_left : n_right;
      _keypoint_buffer.resize(n);
      descriptors.create(n, VSLAM_DESCRIPTOR_BYTES, CV_8UC1);
      int32_t got = 0;
      check(vslam_fpg_get_features(_handle, side, _keypoint_buffer.data(), descriptors.data, n, &got));
      keypoints.resize(n);
      for (int32_t i = 0; i < n; ++i)   // cv::FastFeatureDetector's KeyPoint: size 7, angle -1, octave 0, class_id -1
        keypoints[i] = cv::KeyPoint(_keypoint_buffer[i].x, _keypoint_buffer[i].y, 7.f, -1.f,
                                    _keypoint_buffer[i].response, 0, -1);
    }
    _number_of_detected_keypoints = n_left;                                                          // :101
    double distance = 0;
    check(vslam_fpg_get_detection_stats(_handle, nullptr, nullptr, &distance));
    _current_maximum_descriptor_distance_triangulation = distance;                                   // :109-125
  }
  // :129-132 setFeatures(): the lattices live on the device (row-sorted feature arrays + pruned flags); the host
  // matchers of the base classes stay empty.  Without a new extraction (the tracker's retry, pose_tracker_3d.cpp:320,
  // 402) setFeatures() makes every feature of the frame available again: the device forgets the pruning of the
  // abandoned track() attempt.
  if (!extract_features_) check(vslam_fpg_reset_features(_handle));
  refreshChronometers();
}

void GpuStereoFramePointGenerator::fillPreviousPoint(const FramePoint* point_, vslam_previous_point& out_) {
  const PointCoordinates camera_coordinates(point_->cameraCoordinatesLeft());
  for (int i = 0; i < 3; ++i) out_.camera_left[i] = camera_coordinates(i);
  const bool has_landmark = point_->landmark() != nullptr;
  const PointCoordinates world_coordinates(has_landmark ? point_->landmark()->coordinates() : point_->worldCoordinates());
  for (int i = 0; i < 3; ++i) out_.world[i] = world_coordinates(i);
  std::memcpy(out_.descriptor_left, point_->descriptorLeft().data, VSLAM_DESCRIPTOR_BYTES);
  std::memcpy(out_.descriptor_right, point_->descriptorRight().data, VSLAM_DESCRIPTOR_BYTES);
  out_.epipolar_offset = point_->epipolarOffset();
  out_.has_landmark = has_landmark;
  out_.keypoint_size = point_->keypointLeft().size;
  out_.reserved = 0;
}

void GpuStereoFramePointGenerator::track(Frame* frame_, Frame* frame_previous_,
                                         const TransformMatrix3D& camera_left_previous_in_current_,
                                         FramePointPointerVector& lost_points_, const bool track_by_appearance_) {
  if (!frame_ || !frame_previous_)                                                                   // :468-471
    throw std::runtime_error("StereoFramePointGenerator::track|called with invalid frames");
  FramePointPointerVector& framepoints(frame_->points());
  FramePointPointerVector& framepoints_previous(frame_previous_->points());
  const size_t n_previous = framepoints_previous.size();

  _previous_buffer.resize(n_previous);
  for (size_t i = 0; i < n_previous; ++i) fillPreviousPoint(framepoints_previous[i], _previous_buffer[i]);
  double T[12];   // row-major 3x4 [R|t]
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) T[4 * r + c] = camera_left_previous_in_current_.matrix()(r, c);

  _track_buffer.resize(n_previous);
  _lost_buffer.resize(n_previous);
  int32_t n_tracks = 0, n_lost = 0, n_landmarks = 0;
  double average_distance = 0;
  // :494-672 in two kernels; the matched features are pruned on the device (== :671-672)
  check(vslam_fpg_track(_handle, _previous_buffer.data(), (int32_t)n_previous, T, track_by_appearance_,
                        _projection_tracking_distance_pixels, _maximum_descriptor_distance_tracking, _track_buffer.data(),
                        (int32_t)_track_buffer.size(), &n_tracks, _lost_buffer.data(), &n_lost, &n_landmarks,
                        &average_distance));

  const std::vector<cv::KeyPoint>& kl = frame_->keypointsLeft();
  const std::vector<cv::KeyPoint>& kr = frame_->keypointsRight();
  framepoints.resize(n_tracks);                                                                      // :479, :665
  for (int32_t i = 0; i < n_tracks; ++i) {
    const vslam_track& t = _track_buffer[i];
    const IntensityFeature feature_left(kl[t.index_left], frame_->descriptorsLeft().row(t.index_left), t.index_left);
    const IntensityFeature feature_right(kr[t.index_right], frame_->descriptorsRight().row(t.index_right), t.index_right);
    FramePoint* framepoint = frame_->createFramepoint(&feature_left, &feature_right, t.distance,                // :623-626
                                                      PointCoordinates(t.camera[0], t.camera[1], t.camera[2]),
                                                      framepoints_previous[t.index_previous]);
    framepoint->setEpipolarOffset(t.epipolar_offset);                                                // :627
    framepoint->setProjectionEstimateLeft(cv::Point2f(t.projection_left[0], t.projection_left[1]));  // :631-637
    framepoint->setProjectionEstimateRight(cv::Point2f(t.projection_right[0], t.projection_right[1]));
    framepoint->setProjectionEstimateRightCorrected(
        cv::Point2f(t.projection_right_corrected[0], t.projection_right_corrected[1]));
    framepoints[i] = framepoint;
  }
  // :660-663 a point is lost iff it reached the end of the loop body with !point_previous->next().  next() is set by
  // createFramepoint -- also by an ABANDONED attempt on this frame (the tracker's retry, pose_tracker_3d.cpp:320, 402):
  // the reference then does not report the point as lost, and neither does this
  lost_points_.clear();                                                                              // :482, :666
  lost_points_.reserve(n_lost);
  for (int32_t i = 0; i < n_lost; ++i) {
    FramePoint* point_previous = framepoints_previous[_lost_buffer[i]];
    if (!point_previous->next()) lost_points_.push_back(point_previous);
  }
  _number_of_tracked_landmarks = n_landmarks;                                                        // :653-655
  frame_previous_->setAverageDescriptorDistanceTracking(average_distance);                           // :667-668
}

void GpuStereoFramePointGenerator::recoverPoints(Frame* current_frame_,
                                                 const FramePointPointerVector& lost_points_) const {
  const size_t n_lost = lost_points_.size();
  _previous_buffer.resize(n_lost);
  for (size_t i = 0; i < n_lost; ++i) {
    fillPreviousPoint(lost_points_[i], _previous_buffer[i]);
    if (lost_points_[i]->landmark()) lost_points_[i]->landmark()->incrementNumberOfRecoveries();     // :712
  }
  double W[12];
  const TransformMatrix3D world_to_camera_left = current_frame_->worldToCameraLeft();                // :686-687
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) W[4 * r + c] = world_to_camera_left.matrix()(r, c);
  _recovered_buffer.resize(n_lost);
  int32_t n_recovered = 0;
  check(vslam_fpg_recover_points(_handle, _previous_buffer.data(), (int32_t)n_lost, W, _stereo_parameters->minimum_depth_meters,
                                 _stereo_parameters->maximum_depth_meters, _maximum_descriptor_distance_tracking,
                                 _recovered_buffer.data(), (int32_t)_recovered_buffer.size(), &n_recovered));
  FramePointPointerVector& framepoints(current_frame_->points());
  framepoints.reserve(framepoints.size() + n_recovered);                                             // :698-700, :861
  for (int32_t i = 0; i < n_recovered; ++i) {
    const vslam_recovered_point& r = _recovered_buffer[i];
    FramePoint* point_previous = lost_points_[r.index_lost];
    cv::KeyPoint keypoint_left(point_previous->keypointLeft()), keypoint_right(point_previous->keypointRight());
    keypoint_left.pt = cv::Point2f(r.xl, r.yl);                                                      // :795
    keypoint_right.pt = cv::Point2f(r.xr, r.yr);                                                     // :822
    // FramePoint keeps cv::Mat headers: the descriptors must own their memory (clone), like the ones ORB returns
    const cv::Mat descriptor_left = cv::Mat(1, VSLAM_DESCRIPTOR_BYTES, CV_8UC1, (void*)r.descriptor_left).clone();
    const cv::Mat descriptor_right = cv::Mat(1, VSLAM_DESCRIPTOR_BYTES, CV_8UC1, (void*)r.descriptor_right).clone();
    const IntensityFeature feature_left(keypoint_left, descriptor_left, 0);                          // :846-849
    const IntensityFeature feature_right(keypoint_right, descriptor_right, 0);
    framepoints.push_back(current_frame_->createFramepoint(&feature_left, &feature_right, r.distance,        // :852-856
                                                           PointCoordinates(r.camera[0], r.camera[1], r.camera[2]),
                                                           point_previous));
  }
}

void GpuStereoFramePointGenerator::compute(Frame* frame_) {
  if (!frame_) throw std::runtime_error("StereoFramePointGenerator::compute|called with empty frame");     // :139-142
  FramePointPointerVector& framepoints(frame_->points());

  // what track() pruned (:646-651, :671-672) is already flagged on the device and does not take part in the scan

  // :147-155 points already in the frame (tracked / recovered) pre-load the bin map.  PoseTracker3D::compute may have
  // dropped outliers or appended recovered points since track() (pose_tracker_3d.cpp:437-472, 206): the list is
  // rebuilt from frame->points() as the reference reads it.
  _tracked_buffer.resize(framepoints.size());
  for (size_t i = 0; i < framepoints.size(); ++i) {
    const FramePoint* point = framepoints[i];
    _tracked_buffer[i].row = point->row;
    _tracked_buffer[i].col = point->col;
    _tracked_buffer[i].has_previous = point->previous() != nullptr;
    _tracked_buffer[i].reserved = 0;
    _tracked_buffer[i].disparity = point->disparityPixels();
    _tracked_buffer[i].distance = point->descriptorDistanceTriangulation();
  }

  _framepoint_buffer.resize(frame_->keypointsLeft().size() + framepoints.size() + 1);
  int32_t n_points = 0, n_matches = 0;
  check(vslam_fpg_compute(_handle, _tracked_buffer.data(), (int32_t)_tracked_buffer.size(), _framepoint_buffer.data(),
                          (int32_t)_framepoint_buffer.size(), &n_points, &n_matches));

  // :364-368 + :435-460 : materialise the selected points as FramePoints owned by the frame, in order
  const std::vector<cv::KeyPoint>& kl = frame_->keypointsLeft();
  const std::vector<cv::KeyPoint>& kr = frame_->keypointsRight();
  const size_t number_of_points_tracked = framepoints.size();
  framepoints.reserve(number_of_points_tracked + n_points);
  for (int32_t i = 0; i < n_points; ++i) {
    const vslam_framepoint& r = _framepoint_buffer[i];
    if (r.index_left < 0) {   // an untracked pre-loaded point kept its bin: the reference appends it again (:443-445)
      framepoints.push_back(framepoints[-(r.index_left + 1)]);
      continue;
    }
    const IntensityFeature feature_left(kl[r.index_left], frame_->descriptorsLeft().row(r.index_left), r.index_left);
    const IntensityFeature feature_right(kr[r.index_right], frame_->descriptorsRight().row(r.index_right), r.index_right);
    FramePoint* framepoint = frame_->createFramepoint(&feature_left, &feature_right, r.distance,
                                                      PointCoordinates(r.camera[0], r.camera[1], r.camera[2]));
    framepoint->setEpipolarOffset(r.epipolar_offset);                                                // :368
    framepoints.push_back(framepoint);
  }
  // the matched features are gone from the matchers, as after :419-420 (recoverPoints / the next track() rely on it)
  // NOTE: only the features of the SELECTED points are known here; hosts that need the exact post-compute matcher
  // state call vslam_fpg_get_matches() and prune all n_matches pairs.
  refreshChronometers();
}

// SLAMAssembly::printReport reads the (non-virtual) getTimeConsumptionSeconds_keypoint_detection / _descriptor_extraction /
// _point_triangulation of the base classes (slam_assembly.cpp:690-719; CREATE_CHRONOMETER, definitions.h:144-148): their
// protected counters are kept equal to the device seconds of the corresponding kernels
void GpuStereoFramePointGenerator::refreshChronometers() {
  double detection = 0, extraction = 0, triangulation = 0;
  if (vslam_fpg_get_time_consumption(_handle, &detection, &extraction, &triangulation) != VSLAM_OK) return;
  _time_consumption_seconds_keypoint_detection = detection;
  _time_consumption_seconds_descriptor_extraction = extraction;
  _time_consumption_seconds_point_triangulation = triangulation;
}

void GpuStereoFramePointGenerator::setBriefTests(const int8_t tests_[1024]) {
  std::memcpy(_brief_tests, tests_, sizeof(_brief_tests));
  _brief_tests_set = true;
}

double GpuStereoFramePointGenerator::deviceSecondsKeypointDetection() const {
  double a = 0;
  vslam_fpg_get_time_consumption(_handle, &a, nullptr, nullptr);
  return a;
}
double GpuStereoFramePointGenerator::deviceSecondsDescriptorExtraction() const {
  double a = 0;
  vslam_fpg_get_time_consumption(_handle, nullptr, &a, nullptr);
  return a;
}
double GpuStereoFramePointGenerator::deviceSecondsPointTriangulation() const {
  double a = 0;
  vslam_fpg_get_time_consumption(_handle, nullptr, nullptr, &a);
  return a;
}

}  // namespace proslam
