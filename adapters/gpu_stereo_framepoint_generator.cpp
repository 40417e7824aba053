// gpu_stereo_framepoint_generator.cpp -- see the header.  Every numbered comment cites the reference line whose
// effect the call reproduces (reference src/framepoint_generation/stereo_framepoint_generator.cpp unless noted).
#include "gpu_stereo_framepoint_generator.h"

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
  // the GPU path implements FAST + ORB-256, which is what every stereo configuration of the reference resolves to
  // (parameters.cpp:341 never parses detector_type in stereo mode; BRIEF-256 / ORB-256 fall through to ORB,
  // base_framepoint_generator.cpp:219-224; BRIEF needs opencv_contrib, :185-192)
  if (p->detector_type != "FAST" || p->descriptor_type != "ORB")
    throw std::runtime_error("GpuStereoFramePointGenerator::configure|only FAST + ORB-256 run on the GPU");
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
  // :129-132 the lattice + vectors the inherited track() / recoverPoints() work on
  _feature_matcher_left.setFeatures(frame_->keypointsLeft(), frame_->descriptorsLeft());
  _feature_matcher_right.setFeatures(frame_->keypointsRight(), frame_->descriptorsRight());
}

void GpuStereoFramePointGenerator::compute(Frame* frame_) {
  if (!frame_) throw std::runtime_error("StereoFramePointGenerator::compute|called with empty frame");     // :139-142
  FramePointPointerVector& framepoints(frame_->points());

  // what track() pruned (:646-651, :671-672) must not take part in the scan
  for (int side = 0; side < 2; ++side) {
    const IntensityFeaturePointerVector& remaining =
        side == 0 ? _feature_matcher_left.feature_vector : _feature_matcher_right.feature_vector;
    _keypoint_buffer.resize(remaining.size());
    for (size_t i = 0; i < remaining.size(); ++i) {
      _keypoint_buffer[i].x = remaining[i]->keypoint.pt.x;
      _keypoint_buffer[i].y = remaining[i]->keypoint.pt.y;
      _keypoint_buffer[i].response = remaining[i]->keypoint.response;
    }
    check(vslam_fpg_set_remaining_features(_handle, side, _keypoint_buffer.data(), (int32_t)remaining.size()));
  }

  // :147-155 points already in the frame (tracked / recovered) pre-load the bin map
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
