// oracle/ref/ref_harness.cpp -- see ref_harness.h.  TEST INFRASTRUCTURE ONLY.
// Everything the calls below compute is computed by the reference's own translation units; this file only builds the
// objects the way src/system/slam_assembly.cpp:48-76, :160-200 does, feeds them, and copies results out.  Protected
// state is read through `using` declarations in subclasses that add no behaviour.
#include "ref_harness.h"

#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <string>

#include "aligners/stereouv_aligner.h"
#include "aligners/uvd_aligner.h"
#include "framepoint_generation/stereo_framepoint_generator.h"
#include "position_tracking/pose_tracker_3d.h"
#include "types/landmark.h"
#include "types/world_map.h"
#ifdef VSLAM_REF_WITH_GPU_ADAPTERS
// the drop-in classes of adapters/ (backed by libvslam_b200.so) take the place of the reference's CPU classes
#include "gpu_frame_aligners.h"
#include "gpu_stereo_framepoint_generator.h"
#endif

using namespace proslam;

namespace {

thread_local std::string g_error;

struct Generator : StereoFramePointGenerator {
  using StereoFramePointGenerator::StereoFramePointGenerator;
  using StereoFramePointGenerator::_current_maximum_descriptor_distance_triangulation;
  using StereoFramePointGenerator::_detectors;
  using StereoFramePointGenerator::_feature_matcher_left;
  using StereoFramePointGenerator::_feature_matcher_right;
  using StereoFramePointGenerator::_number_of_tracked_landmarks;
};

template <class Base, int Dim>
struct AlignerAccess : Base {
  using Base::Base;
  using Base::_b;
  using Base::_camera_calibration_matrix;
  using Base::_errors;
  using Base::_fixed;
  using Base::_H;
  using Base::_information_matrix;
  using Base::_information_matrix_vector;
  using Base::_inliers;
  using Base::_minimum_reliable_depth_meters;
  using Base::_moving;
  using Base::_number_of_cols_image;
  using Base::_number_of_inliers;
  using Base::_number_of_measurements;
  using Base::_number_of_outliers;
  using Base::_number_of_rows_image;
  using Base::_frame_current;
  using Base::_frame_previous;
  using Base::_previous_to_current;
  using Base::_total_error;
  using Base::_weights_translation;
  int rounds = 0;
  void oneRound(const bool& ignore_outliers_) override {
    ++rounds;
    Base::oneRound(ignore_outliers_);
  }
};
struct UV : AlignerAccess<StereoUVAligner, 4> {
  using AlignerAccess<StereoUVAligner, 4>::AlignerAccess;
  using StereoUVAligner::_offset_camera_right;
};
typedef AlignerAccess<UVDAligner, 3> UVD;
#ifdef VSLAM_REF_WITH_GPU_ADAPTERS
struct GpuUV : AlignerAccess<GpuStereoUVAligner, 4> {
  using AlignerAccess<GpuStereoUVAligner, 4>::AlignerAccess;
  using StereoUVAligner::_offset_camera_right;
};
#else
struct GpuUV;
#endif

struct Tracker : PoseTracker3D {
  using PoseTracker3D::PoseTracker3D;
  using PoseTracker3D::_number_of_active_landmarks;
  using PoseTracker3D::_number_of_tracked_points;
  using PoseTracker3D::_status;
};

TransformMatrix3D to_transform(const double t[12]) {
  TransformMatrix3D r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) r.matrix()(i, j) = t[4 * i + j];
  return r;
}
void from_transform(const TransformMatrix3D& t, double out[12]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) out[4 * i + j] = t.matrix()(i, j);
}

}  // namespace

struct ref_session {
  ParameterCollection* parameters = nullptr;
  Camera* camera_left = nullptr;
  Camera* camera_right = nullptr;
  WorldMap* world = nullptr;
  Tracker* tracker = nullptr;         // owns generator and uv (pose_tracker_3d.cpp:27-28)
  Generator* generator = nullptr;     // CPU mode only (nullptr when the GPU adapters are in place)
  StereoFramePointGenerator* base_generator = nullptr;   // either mode
  BaseFrameAligner* base_uv = nullptr;
  bool gpu = false;
  GpuUV* gpu_uv = nullptr;            // GPU mode: the tracker's aligner
  UV* uv = nullptr;                   // CPU mode: the tracker's aligner
  UVD* uvd = nullptr;
  AlignerParameters* uvd_parameters = nullptr;
  FramePointPointerVector lost;
  Frame* loaded_frame = nullptr;      // holder of dummy points for ref_aligner_load + converge()'s visualisation loop
  std::vector<IntensityFeature*> dummy_features;
  int rows = 0, cols = 0;
  double K[9];
  double bx = 0;
};

#define REF_TRY try {
#define REF_CATCH(fail)                                       \
  }                                                           \
  catch (const std::exception& e) {                           \
    g_error = e.what();                                       \
    return fail;                                              \
  }                                                           \
  catch (...) {                                               \
    g_error = "unknown exception";                            \
    return fail;                                              \
  }

static_assert(sizeof(ref_point) == 248, "oracle/ref.py POINT mirrors ref_point");

extern "C" const char* ref_last_error(void) { return g_error.c_str(); }

extern "C" ref_session* ref_open(const char* yaml, int rows, int cols, const double K[9], double bx) {
  REF_TRY
  // the reference prints unguarded diagnostics to stdout for every detection (base_framepoint_generator.cpp:369-379)
  static bool silenced = false;
  if (!silenced) {
    std::cout.setstate(std::ios::failbit);
    silenced = true;
  }
  ref_session* s = new ref_session();
  s->rows = rows;
  s->cols = cols;
  std::memcpy(s->K, K, sizeof(s->K));
  s->bx = bx;
  s->parameters = new ParameterCollection();
  if (yaml && *yaml) {
    std::ifstream probe(yaml);
    if (!probe) {            // parseFromFile only logs a YAML::BadFile (parameters.cpp:438); tests want to know
      g_error = std::string("cannot open ") + yaml;
      delete s->parameters;
      delete s;
      return nullptr;
    }
    s->parameters->parseFromFile(yaml);
  } else {
    s->parameters->setMode(CommandLineParameters::TrackerMode::RGB_STEREO);
  }
  return s;
  REF_CATCH(nullptr)
}

extern "C" void ref_close(ref_session* s) {
  if (!s) return;
  try {
    for (IntensityFeature* f : s->dummy_features) delete f;
    delete s->tracker;       // deletes generator and uv
    delete s->uvd;
    delete s->uvd_parameters;
    delete s->world;         // deletes frames (and their points), landmarks
    delete s->camera_left;
    delete s->camera_right;
    delete s->parameters;
  } catch (...) {
  }
  delete s;
}

static void copy_name(char dst[32], const std::string& src) {
  std::memset(dst, 0, 32);
  std::strncpy(dst, src.c_str(), 31);
}

extern "C" int ref_get_parameters(ref_session* s, ref_parameters* o) {
  REF_TRY
  const StereoFramePointGeneratorParameters* g = s->parameters->stereo_framepoint_generator_parameters;
  const PoseTracker3DParameters* t = s->parameters->tracker_parameters;
  const AlignerParameters* a = t->aligner;
  std::memset(o, 0, sizeof(*o));
  copy_name(o->detector_type, g->detector_type);
  copy_name(o->descriptor_type, g->descriptor_type);
  o->target_number_of_keypoints_tolerance = g->target_number_of_keypoints_tolerance;
  o->detector_threshold_minimum = g->detector_threshold_minimum;
  o->detector_threshold_maximum = g->detector_threshold_maximum;
  o->detector_threshold_maximum_change = g->detector_threshold_maximum_change;
  o->number_of_detectors_vertical = g->number_of_detectors_vertical;
  o->number_of_detectors_horizontal = g->number_of_detectors_horizontal;
  o->minimum_projection_tracking_distance_pixels = g->minimum_projection_tracking_distance_pixels;
  o->maximum_projection_tracking_distance_pixels = g->maximum_projection_tracking_distance_pixels;
  o->minimum_descriptor_distance_tracking = g->minimum_descriptor_distance_tracking;
  o->maximum_descriptor_distance_tracking = g->maximum_descriptor_distance_tracking;
  o->maximum_reliable_depth_meters = g->maximum_reliable_depth_meters;
  o->maximum_depth_meters = g->maximum_depth_meters;
  o->minimum_depth_meters = g->minimum_depth_meters;
  o->enable_keypoint_binning = g->enable_keypoint_binning;
  o->bin_size_pixels = g->bin_size_pixels;
  o->maximum_matching_distance_triangulation = g->maximum_matching_distance_triangulation;
  o->minimum_disparity_pixels = g->minimum_disparity_pixels;
  o->maximum_epipolar_search_offset_pixels = g->maximum_epipolar_search_offset_pixels;
  o->use_matches = g->use_matches;
  o->error_delta_for_convergence = a->error_delta_for_convergence;
  o->maximum_error_kernel = a->maximum_error_kernel;
  o->damping = a->damping;
  o->maximum_number_of_iterations = a->maximum_number_of_iterations;
  o->minimum_number_of_inliers = a->minimum_number_of_inliers;
  o->minimum_inlier_ratio = a->minimum_inlier_ratio;
  o->enable_inverse_depth_as_information = a->enable_inverse_depth_as_information;
  o->minimum_track_length_for_landmark_creation = t->minimum_track_length_for_landmark_creation;
  o->minimum_number_of_landmarks_to_track = t->minimum_number_of_landmarks_to_track;
  o->tunnel_vision_ratio = t->tunnel_vision_ratio;
  o->good_tracking_ratio = t->good_tracking_ratio;
  o->maximum_number_of_landmark_recoveries = t->maximum_number_of_landmark_recoveries;
  o->enable_landmark_recovery = t->enable_landmark_recovery;
  o->motion_model = static_cast<int32_t>(t->motion_model);
  o->minimum_delta_angular_for_movement = t->minimum_delta_angular_for_movement;
  o->minimum_delta_translational_for_movement = t->minimum_delta_translational_for_movement;
  o->maximum_error_squared_meters = s->parameters->world_map_parameters->landmark->maximum_error_squared_meters;
  return 0;
  REF_CATCH(-1)
}

extern "C" int ref_set_parameters(ref_session* s, const ref_parameters* i) {
  REF_TRY
  if (s->tracker) throw std::runtime_error("ref_set_parameters after ref_configure");
  StereoFramePointGeneratorParameters* g = s->parameters->stereo_framepoint_generator_parameters;
  PoseTracker3DParameters* t = s->parameters->tracker_parameters;
  AlignerParameters* a = t->aligner;
  g->detector_type = i->detector_type;
  g->descriptor_type = i->descriptor_type;
  g->target_number_of_keypoints_tolerance = i->target_number_of_keypoints_tolerance;
  g->detector_threshold_minimum = static_cast<uint32_t>(i->detector_threshold_minimum);
  g->detector_threshold_maximum = static_cast<uint32_t>(i->detector_threshold_maximum);
  g->detector_threshold_maximum_change = i->detector_threshold_maximum_change;
  g->number_of_detectors_vertical = i->number_of_detectors_vertical;
  g->number_of_detectors_horizontal = i->number_of_detectors_horizontal;
  g->minimum_projection_tracking_distance_pixels = i->minimum_projection_tracking_distance_pixels;
  g->maximum_projection_tracking_distance_pixels = i->maximum_projection_tracking_distance_pixels;
  g->minimum_descriptor_distance_tracking = i->minimum_descriptor_distance_tracking;
  g->maximum_descriptor_distance_tracking = i->maximum_descriptor_distance_tracking;
  g->maximum_reliable_depth_meters = i->maximum_reliable_depth_meters;
  g->maximum_depth_meters = i->maximum_depth_meters;
  g->minimum_depth_meters = i->minimum_depth_meters;
  g->enable_keypoint_binning = i->enable_keypoint_binning != 0;
  g->bin_size_pixels = i->bin_size_pixels;
  g->maximum_matching_distance_triangulation = i->maximum_matching_distance_triangulation;
  g->minimum_disparity_pixels = i->minimum_disparity_pixels;
  g->maximum_epipolar_search_offset_pixels = i->maximum_epipolar_search_offset_pixels;
  g->use_matches = i->use_matches != 0;
  a->error_delta_for_convergence = i->error_delta_for_convergence;
  a->maximum_error_kernel = i->maximum_error_kernel;
  a->damping = i->damping;
  a->maximum_number_of_iterations = i->maximum_number_of_iterations;
  a->minimum_number_of_inliers = i->minimum_number_of_inliers;
  a->minimum_inlier_ratio = i->minimum_inlier_ratio;
  a->enable_inverse_depth_as_information = i->enable_inverse_depth_as_information != 0;
  t->minimum_track_length_for_landmark_creation = i->minimum_track_length_for_landmark_creation;
  t->minimum_number_of_landmarks_to_track = i->minimum_number_of_landmarks_to_track;
  t->tunnel_vision_ratio = i->tunnel_vision_ratio;
  t->good_tracking_ratio = i->good_tracking_ratio;
  t->maximum_number_of_landmark_recoveries = i->maximum_number_of_landmark_recoveries;
  t->enable_landmark_recovery = i->enable_landmark_recovery != 0;
  t->motion_model = static_cast<Parameters::MotionModel>(i->motion_model);
  t->minimum_delta_angular_for_movement = i->minimum_delta_angular_for_movement;
  t->minimum_delta_translational_for_movement = i->minimum_delta_translational_for_movement;
  s->parameters->world_map_parameters->landmark->maximum_error_squared_meters = i->maximum_error_squared_meters;
  return 0;
  REF_CATCH(-1)
}

static int configure(ref_session* s, bool gpu) {
  REF_TRY
  if (s->tracker) throw std::runtime_error("ref_configure called twice");
  s->gpu = gpu;
  // counters are process-wide statics (slam_assembly.cpp:28-32)
  Frame::reset();
  FramePoint::reset();
  LocalMap::reset();
  Landmark::reset();
  // cameras as SLAMAssembly::loadCamerasFromMessageFile leaves them (slam_assembly.cpp:160-200): P = K [I | 0] and
  // K [I | t]; the right camera's homogeneous baseline is its projection matrix' last column
  CameraMatrix K;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) K(i, j) = s->K[3 * i + j];
  s->camera_left = new Camera(s->rows, s->cols, K);
  s->camera_right = new Camera(s->rows, s->cols, K);
  ProjectionMatrix P = ProjectionMatrix::Zero();
  P.block<3, 3>(0, 0) = K;
  s->camera_left->setProjectionMatrix(P);
  P(0, 3) = s->bx;
  s->camera_right->setProjectionMatrix(P);
  s->camera_right->setBaselineHomogeneous(P.col(3));

  s->world = new WorldMap(s->parameters->world_map_parameters);
  s->tracker = new Tracker(s->parameters->tracker_parameters);
  s->tracker->setWorldMap(s->world);
  s->tracker->setCameraLeft(s->camera_left);
  s->tracker->setCameraSecondary(s->camera_right);

  // SLAMAssembly::_createStereoTracker (slam_assembly.cpp:48-76)
  s->camera_left->setCameraMatrix(s->camera_left->projectionMatrix().block<3, 3>(0, 0));
  s->camera_right->setCameraMatrix(s->camera_left->cameraMatrix());
  if (!gpu) {
    s->generator = new Generator(s->parameters->stereo_framepoint_generator_parameters);
    s->base_generator = s->generator;
    s->uv = new UV(s->parameters->tracker_parameters->aligner);
    s->base_uv = s->uv;
  } else {
#ifdef VSLAM_REF_WITH_GPU_ADAPTERS
    // the two `new` expressions of slam_assembly.cpp:62, :68 name the GPU classes instead
    s->base_generator = new GpuStereoFramePointGenerator(s->parameters->stereo_framepoint_generator_parameters);
    s->gpu_uv = new GpuUV(s->parameters->tracker_parameters->aligner);
    s->base_uv = s->gpu_uv;
#else
    throw std::runtime_error("this build of oracle/_ref has no GPU adapters (make _ref_gpu)");
#endif
  }
  s->base_generator->setCameraLeft(s->camera_left);
  s->base_generator->setCameraRight(s->camera_right);
  s->base_generator->configure();
  s->base_uv->setMaximumReliableDepthMeters(s->parameters->stereo_framepoint_generator_parameters->maximum_reliable_depth_meters);
  s->base_uv->setMinimumReliableDepthMeters(s->parameters->stereo_framepoint_generator_parameters->minimum_depth_meters);
  s->base_uv->configure();
  s->tracker->setFramePointGenerator(s->base_generator);
  s->tracker->setAligner(s->base_uv);
  s->tracker->configure();

  // the depth aligner (slam_assembly.cpp:88-92), on a copy of the same AlignerParameters
  s->uvd_parameters = new AlignerParameters(*s->parameters->tracker_parameters->aligner);
  s->uvd = new UVD(s->uvd_parameters);
  s->uvd->setMaximumReliableDepthMeters(s->parameters->stereo_framepoint_generator_parameters->maximum_reliable_depth_meters);
  s->uvd->setMinimumReliableDepthMeters(s->parameters->stereo_framepoint_generator_parameters->minimum_depth_meters);
  s->uvd->configure();
  return 0;
  REF_CATCH(-1)
}

extern "C" int ref_configure(ref_session* s) { return configure(s, false); }
extern "C" int ref_configure_gpu(ref_session* s) { return configure(s, true); }
extern "C" int ref_has_gpu_adapters(void) {
#ifdef VSLAM_REF_WITH_GPU_ADAPTERS
  return 1;
#else
  return 0;
#endif
}

static Generator* cpu_generator(ref_session* s) {
  if (!s->generator) throw std::runtime_error("this call reads the CPU generator's internals (not available with the GPU adapters)");
  return s->generator;
}

// ---- generator ---------------------------------------------------------------------------------------------------
static cv::Mat copy_image(const uint8_t* data, int rows, int cols, int stride) {
  cv::Mat m(rows, cols, CV_8UC1);
  for (int y = 0; y < rows; ++y) std::memcpy(m.ptr(y), data + (size_t)y * stride, cols);
  return m;
}

extern "C" int ref_fpg_initialize(ref_session* s, const uint8_t* left, const uint8_t* right, int stride, int status) {
  REF_TRY
  // the frame set-up of PoseTracker3D::compute (pose_tracker_3d.cpp:69-80)
  Frame* frame = s->world->createFrame();
  frame->setCameraLeft(s->camera_left);
  frame->setIntensityImageLeft(copy_image(left, s->rows, s->cols, stride));
  frame->setCameraRight(s->camera_right);
  frame->setIntensityImageRight(copy_image(right, s->rows, s->cols, stride));
  frame->setStatus(status ? Frame::Tracking : Frame::Localizing);
  s->base_generator->initialize(frame);
  s->lost.clear();
  return 0;
  REF_CATCH(-1)
}

extern "C" int ref_fpg_reinitialize(ref_session* s) {
  REF_TRY
  s->base_generator->initialize(s->world->currentFrame(), false);
  return 0;
  REF_CATCH(-1)
}

extern "C" int ref_fpg_features(ref_session* s, int side, float* xyr, uint8_t* desc, int capacity) {
  REF_TRY
  Frame* f = s->world->currentFrame();
  const std::vector<cv::KeyPoint>& k = side ? f->keypointsRight() : f->keypointsLeft();
  const cv::Mat& d = side ? f->descriptorsRight() : f->descriptorsLeft();
  const int n = (int)k.size();
  if (n > capacity) throw std::runtime_error("ref_fpg_features: capacity");
  if (n && d.rows != n) throw std::runtime_error("ref_fpg_features: keypoints and descriptors disagree");
  for (int i = 0; i < n; ++i) {
    if (xyr) {
      xyr[3 * i] = k[i].pt.x;
      xyr[3 * i + 1] = k[i].pt.y;
      xyr[3 * i + 2] = k[i].response;
    }
    if (desc) std::memcpy(desc + 32 * (size_t)i, d.ptr(i), 32);
  }
  return n;
  REF_CATCH(-1)
}

extern "C" int ref_fpg_remaining(ref_session* s, int side, float* xy, int capacity) {
  REF_TRY
  const IntensityFeaturePointerVector& v =
      side ? cpu_generator(s)->_feature_matcher_right.feature_vector : cpu_generator(s)->_feature_matcher_left.feature_vector;
  const int n = (int)v.size();
  if (n > capacity) throw std::runtime_error("ref_fpg_remaining: capacity");
  for (int i = 0; i < n; ++i) {
    xy[2 * i] = v[i]->keypoint.pt.x;
    xy[2 * i + 1] = v[i]->keypoint.pt.y;
  }
  return n;
  REF_CATCH(-1)
}

extern "C" int ref_fpg_thresholds(ref_session* s, double* out, int capacity) {
  REF_TRY
  const StereoFramePointGeneratorParameters* g = s->parameters->stereo_framepoint_generator_parameters;
  const int n = (int)(g->number_of_detectors_vertical * g->number_of_detectors_horizontal);
  if (n > capacity) throw std::runtime_error("ref_fpg_thresholds: capacity");
  for (uint32_t r = 0; r < g->number_of_detectors_vertical; ++r)
    for (uint32_t c = 0; c < g->number_of_detectors_horizontal; ++c)
      out[r * g->number_of_detectors_horizontal + c] = cpu_generator(s)->_detectors[r][c]->getThreshold();
  return n;
  REF_CATCH(-1)
}

extern "C" double ref_fpg_triangulation_distance(ref_session* s) {
  return s->generator ? s->generator->_current_maximum_descriptor_distance_triangulation : -1.0;
}
extern "C" int ref_fpg_target_number_of_keypoints(ref_session* s) { return (int)s->base_generator->targetNumberOfKeypoints(); }

extern "C" int ref_fpg_set_tracking(ref_session* s, int distance_pixels, double maximum_descriptor_distance) {
  REF_TRY
  s->base_generator->setProjectionTrackingDistancePixels(distance_pixels);
  s->base_generator->setMaximumDescriptorDistanceTracking(maximum_descriptor_distance);
  return 0;
  REF_CATCH(-1)
}

static int index_in(const FramePointPointerVector& v, const FramePoint* p) {
  for (size_t i = 0; i < v.size(); ++i)
    if (v[i] == p) return (int)i;
  return -1;
}

extern "C" int ref_fpg_track(ref_session* s, const double T[12], int by_appearance, int32_t* lost, int* n_lost,
                             int* number_of_tracked_landmarks, double* average_descriptor_distance) {
  REF_TRY
  Frame* current = s->world->currentFrame();
  Frame* previous = current->previous();
  if (!previous) throw std::runtime_error("ref_fpg_track: no previous frame");
  // index of every previous point BEFORE the call (track() does not reorder previous->points())
  std::map<const FramePoint*, int> position;
  for (size_t i = 0; i < previous->points().size(); ++i) position[previous->points()[i]] = (int)i;
  s->lost.clear();
  s->base_generator->track(current, previous, to_transform(T), s->lost, by_appearance != 0);
  if (n_lost) *n_lost = (int)s->lost.size();
  if (lost)
    for (size_t i = 0; i < s->lost.size(); ++i) lost[i] = position.at(s->lost[i]);
  if (number_of_tracked_landmarks) *number_of_tracked_landmarks = (int)s->base_generator->numberOfTrackedLandmarks();
  if (average_descriptor_distance) *average_descriptor_distance = previous->averageDescriptorDistanceTracking();   // set on the PREVIOUS frame (:664)
  return (int)current->points().size();
  REF_CATCH(-1)
}

extern "C" int ref_fpg_recover(ref_session* s) {
  REF_TRY
  Frame* current = s->world->currentFrame();
  const size_t before = current->points().size();
  s->base_generator->recoverPoints(current, s->lost);
  return (int)(current->points().size() - before);
  REF_CATCH(-1)
}

extern "C" int ref_fpg_compute(ref_session* s) {
  REF_TRY
  s->base_generator->compute(s->world->currentFrame());
  return (int)s->world->currentFrame()->points().size();
  REF_CATCH(-1)
}

extern "C" int ref_frame_points(ref_session* s, int which, ref_point* out, int capacity) {
  REF_TRY
  Frame* f = s->world->currentFrame();
  if (which == 1) f = f ? f->previous() : nullptr;
  if (!f) throw std::runtime_error("ref_frame_points: no such frame");
  const FramePointPointerVector& pts = f->points();
  if ((int)pts.size() > capacity) throw std::runtime_error("ref_frame_points: capacity");
  std::map<const FramePoint*, int> position;
  if (f->previous())
    for (size_t i = 0; i < f->previous()->points().size(); ++i) position[f->previous()->points()[i]] = (int)i;
  for (size_t i = 0; i < pts.size(); ++i) {
    const FramePoint* p = pts[i];
    ref_point& o = out[i];
    std::memset(&o, 0, sizeof(o));
    o.xl = p->keypointLeft().pt.x; o.yl = p->keypointLeft().pt.y;
    o.xr = p->keypointRight().pt.x; o.yr = p->keypointRight().pt.y;
    o.row = p->row; o.col = p->col;
    o.epipolar_offset = p->epipolarOffset();
    o.index_previous = -1;
    if (p->previous()) {
      auto it = position.find(p->previous());
      o.index_previous = it == position.end() ? -2 : it->second;     // -2: linked, but no longer in points()
    }
    o.disparity = p->disparityPixels();
    o.distance = p->descriptorDistanceTriangulation();
    for (int k = 0; k < 3; ++k) {
      o.cam[k] = p->cameraCoordinatesLeft()(k);
      o.robot[k] = p->robotCoordinates()(k);
      o.world[k] = p->worldCoordinates()(k);
    }
    o.projection_left[0] = p->projectionEstimateLeft().x; o.projection_left[1] = p->projectionEstimateLeft().y;
    o.projection_right[0] = p->projectionEstimateRight().x; o.projection_right[1] = p->projectionEstimateRight().y;
    o.projection_right_corrected[0] = p->projectionEstimateRightCorrected().x;
    o.projection_right_corrected[1] = p->projectionEstimateRightCorrected().y;
    o.has_landmark = p->landmark() != nullptr;
    o.track_length = p->trackLength();
    if (p->landmark()) {
      for (int k = 0; k < 3; ++k) o.landmark_world[k] = p->landmark()->coordinates()(k);
      o.landmark_updates = p->landmark()->numberOfUpdates();
    }
    if (!p->descriptorLeft().empty()) std::memcpy(o.desc_left, p->descriptorLeft().ptr(0), 32);
    if (!p->descriptorRight().empty()) std::memcpy(o.desc_right, p->descriptorRight().ptr(0), 32);
  }
  return (int)pts.size();
  REF_CATCH(-1)
}

extern "C" int ref_frame_set_pose(ref_session* s, const double robot_to_world[12]) {
  REF_TRY
  s->world->currentFrame()->setRobotToWorld(to_transform(robot_to_world));
  s->world->setRobotToWorld(s->world->currentFrame()->robotToWorld());
  return 0;
  REF_CATCH(-1)
}

extern "C" int ref_frame_make_landmarks(ref_session* s, int every_nth) {
  REF_TRY
  // what PoseTracker3D::_updatePoints does for a mature point (pose_tracker_3d.cpp:486-519), applied to every n-th
  // point of the current frame that has a previous point
  Frame* f = s->world->currentFrame();
  int made = 0, k = 0;
  for (FramePoint* point : f->points()) {
    point->setWorldCoordinates(f->robotToWorld() * point->robotCoordinates());
    if (!point->previous()) continue;
    if (every_nth > 1 && (k++ % every_nth) != 0) continue;
    Landmark* landmark = point->origin()->landmark();
    if (!landmark) landmark = s->world->createLandmark(point);
    else landmark->update(point);
    point->setCameraCoordinatesLeftLandmark(f->worldToCameraLeft() * landmark->coordinates());
    ++made;
  }
  return made;
  REF_CATCH(-1)
}

extern "C" double ref_fpg_seconds(ref_session* s, int which) {
  switch (which) {
    case 0: return s->base_generator->getTimeConsumptionSeconds_keypoint_detection();
    case 1: return s->base_generator->getTimeConsumptionSeconds_descriptor_extraction();
    default: return s->base_generator->getTimeConsumptionSeconds_point_triangulation();
  }
}

// ---- aligners ------------------------------------------------------------------------------------------------------
// converge() ends with a visualisation loop over _frame_current->points() (stereouv_aligner.cpp:258-263): a loaded problem
// needs a frame that holds as many points.
static Frame* frame_with_points(ref_session* s, int n) {
  Frame* f = s->world->createFrame();
  f->setCameraLeft(s->camera_left);
  f->setCameraRight(s->camera_right);
  IntensityFeature* feature = new IntensityFeature(cv::KeyPoint(cv::Point2f(0, 0), 7), cv::Mat(), 0);
  s->dummy_features.push_back(feature);
  f->points().reserve(n);
  for (int i = 0; i < n; ++i) f->points().push_back(f->createFramepoint(feature, feature, 0, PointCoordinates(0, 0, 1)));
  return f;
}

template <class A>
static void load_common(ref_session* s, A* a, int n, const double* moving, const double* wt, double min_depth) {
  a->_frame_previous = nullptr;
  a->_frame_current = frame_with_points(s, n);
  a->_number_of_measurements = n;
  a->_errors.assign(n, 0);
  a->_inliers.assign(n, false);
  a->_information_matrix_vector.resize(n);
  a->_weights_translation.assign(wt, wt + n);
  a->_moving.resize(n);
  a->_fixed.resize(n);
  for (int u = 0; u < n; ++u) a->_moving[u] = Vector3(moving[3 * u], moving[3 * u + 1], moving[3 * u + 2]);
  a->_camera_calibration_matrix = s->camera_left->cameraMatrix();
  a->_number_of_rows_image = s->rows;
  a->_number_of_cols_image = s->cols;
  a->_minimum_reliable_depth_meters = min_depth;
  a->_previous_to_current.setIdentity();
  a->rounds = 0;
}

extern "C" int ref_aligner_load(ref_session* s, int kind, int n, const double* moving, const double* fixed,
                                const double* omega, const double* wt, const double baseline[3], double min_depth) {
  REF_TRY
  if (kind == 0 && s->gpu)
    throw std::runtime_error("ref_aligner_load: explicit correspondences go through the C ABI in GPU mode (tests/test_gpu_aligner.py)");
  if (kind == 0) {
    load_common(s, s->uv, n, moving, wt, min_depth);
    for (int u = 0; u < n; ++u) {
      s->uv->_fixed[u] = Vector4(fixed[4 * u], fixed[4 * u + 1], fixed[4 * u + 2], fixed[4 * u + 3]);
      s->uv->_information_matrix_vector[u].setIdentity();          // stereouv_aligner.cpp:29, :50
      s->uv->_information_matrix_vector[u] *= omega[u];
    }
    s->uv->_offset_camera_right = Vector3(baseline[0], baseline[1], baseline[2]);
  } else {
    load_common(s, s->uvd, n, moving, wt, min_depth);
    for (int u = 0; u < n; ++u) {
      s->uvd->_fixed[u] = Vector3(fixed[3 * u], fixed[3 * u + 1], fixed[3 * u + 2]);
      s->uvd->_information_matrix_vector[u].setZero();             // diag(w, w, w_depth): uvd_aligner.cpp:30-61
      s->uvd->_information_matrix_vector[u](0, 0) = omega[2 * u];
      s->uvd->_information_matrix_vector[u](1, 1) = omega[2 * u];
      s->uvd->_information_matrix_vector[u](2, 2) = omega[2 * u + 1];
    }
  }
  return 0;
  REF_CATCH(-1)
}

#ifdef VSLAM_REF_WITH_GPU_ADAPTERS
#define WITH_GPU_ALIGNER(stmt)    \
  {                               \
    GpuUV* a = s->gpu_uv;         \
    stmt;                         \
  }
#else
#define WITH_GPU_ALIGNER(stmt) throw std::runtime_error("no GPU adapters in this build");
#endif
#define WITH_ALIGNER(stmt)        \
  if (kind == 0 && s->gpu) {      \
    WITH_GPU_ALIGNER(stmt)        \
  } else if (kind == 0) {         \
    UV* a = s->uv;                \
    stmt;                         \
  } else {                        \
    UVD* a = s->uvd;              \
    stmt;                         \
  }

extern "C" int ref_aligner_set_pose(ref_session* s, int kind, const double T[12]) {
  REF_TRY
  WITH_ALIGNER(a->_previous_to_current = to_transform(T))
  return 0;
  REF_CATCH(-1)
}
extern "C" int ref_aligner_linearize(ref_session* s, int kind, int ignore_outliers) {
  REF_TRY
  WITH_ALIGNER(a->linearize(ignore_outliers != 0))
  return 0;
  REF_CATCH(-1)
}
extern "C" int ref_aligner_one_round(ref_session* s, int kind, int ignore_outliers) {
  REF_TRY
  WITH_ALIGNER(a->oneRound(ignore_outliers != 0))
  return 0;
  REF_CATCH(-1)
}
extern "C" int ref_aligner_converge(ref_session* s, int kind) {
  REF_TRY
  WITH_ALIGNER(a->converge(); return a->hasSystemConverged() ? 1 : 0)
  REF_CATCH(-1)
}

template <class A>
static void dump_state(A* a, double H[36], double b[6], double* total_error, int* inliers, int* outliers, double T[12],
                       double* errors, uint8_t* flags, double information[36]) {
  if (H)
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) H[6 * i + j] = a->_H(i, j);
  if (b)
    for (int i = 0; i < 6; ++i) b[i] = a->_b(i);
  if (total_error) *total_error = a->_total_error;
  if (inliers) *inliers = (int)a->_number_of_inliers;
  if (outliers) *outliers = (int)a->_number_of_outliers;
  if (T) from_transform(a->_previous_to_current, T);
  const size_t n = a->_number_of_measurements;
  if (errors)
    for (size_t u = 0; u < n; ++u) errors[u] = a->_errors[u];
  if (flags)
    for (size_t u = 0; u < n; ++u) flags[u] = a->_inliers[u] ? 1 : 0;
  if (information)
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) information[6 * i + j] = a->_information_matrix(i, j);
}

extern "C" int ref_aligner_state(ref_session* s, int kind, double H[36], double b[6], double* total_error, int* inliers,
                                 int* outliers, double T[12], double* errors, uint8_t* flags, double information[36]) {
  REF_TRY
  WITH_ALIGNER(dump_state(a, H, b, total_error, inliers, outliers, T, errors, flags, information))
  return 0;
  REF_CATCH(-1)
}

extern "C" int ref_aligner_initialize_frames(ref_session* s, int kind, const double T0[12], int enable_inverse_depth) {
  REF_TRY
  Frame* current = s->world->currentFrame();
  if (!current || !current->previous()) throw std::runtime_error("ref_aligner_initialize_frames: two frames needed");
  WITH_ALIGNER(a->parameters()->enable_inverse_depth_as_information = enable_inverse_depth != 0;   // pose_tracker_3d.cpp:124, :355
               a->rounds = 0;
               a->initialize(current->previous(), current, to_transform(T0));
               return (int)a->_number_of_measurements)
  REF_CATCH(-1)
}

template <class A>
static int packed_uv(A* a, double* moving, double* fixed, double* omega, double* wt) {
  for (size_t u = 0; u < a->_number_of_measurements; ++u) {
    for (int k = 0; k < 3; ++k) moving[3 * u + k] = a->_moving[u](k);
    for (int k = 0; k < 4; ++k) fixed[4 * u + k] = a->_fixed[u](k);
    omega[u] = a->_information_matrix_vector[u](0, 0);
    wt[u] = a->_weights_translation[u];
  }
  return (int)a->_number_of_measurements;
}

extern "C" int ref_aligner_packed(ref_session* s, int kind, double* moving, double* fixed, double* omega, double* wt) {
  REF_TRY
  if (kind == 0) {
    WITH_ALIGNER(return packed_uv(a, moving, fixed, omega, wt))
  }
  UVD* a = s->uvd;
  for (size_t u = 0; u < a->_number_of_measurements; ++u) {
    for (int k = 0; k < 3; ++k) moving[3 * u + k] = a->_moving[u](k);
    for (int k = 0; k < 3; ++k) fixed[3 * u + k] = a->_fixed[u](k);
    omega[2 * u] = a->_information_matrix_vector[u](0, 0);
    omega[2 * u + 1] = a->_information_matrix_vector[u](2, 2);
    wt[u] = a->_weights_translation[u];
  }
  return (int)a->_number_of_measurements;
  REF_CATCH(-1)
}
extern "C" int ref_aligner_count(ref_session* s, int kind) {
  WITH_ALIGNER(return (int)a->_number_of_measurements)
}
extern "C" int ref_aligner_rounds(ref_session* s, int kind) {
  WITH_ALIGNER(return a->rounds)
}

// ---- the batched first-frame workload of bench.py (BASELINE configs[2]) on the reference's classes ------------------------
extern "C" int ref_reset(ref_session* s) {
  REF_TRY
  // a fresh sequence: no frames, no landmarks, detector thresholds back at their configured minimum
  // (base_framepoint_generator.cpp:244-251), tracker state as constructed
  s->world->clear();
  s->world->setCurrentFrame(nullptr);
  s->world->setPreviousFrame(nullptr);
  s->lost.clear();
  if (s->generator) {
    const StereoFramePointGeneratorParameters* g = s->parameters->stereo_framepoint_generator_parameters;
    for (uint32_t r = 0; r < g->number_of_detectors_vertical; ++r)
      for (uint32_t c = 0; c < g->number_of_detectors_horizontal; ++c)
        s->generator->_detectors[r][c]->setThreshold(g->detector_threshold_minimum);
  }
  return 0;
  REF_CATCH(-1)
}

extern "C" int ref_first_frame(ref_session* s, const uint8_t* left, const uint8_t* right, int stride, int rounds,
                               const double T[12], double* seconds_pose_optimization) {
  REF_TRY
  if (ref_reset(s) < 0 || ref_fpg_initialize(s, left, right, stride, 0) < 0) return -1;
  s->base_generator->compute(s->world->currentFrame());
  // StereoUVAligner over the frame's own new points at a small prior error (previous frame == current frame, no
  // landmarks: the first-frame analogue of pose_tracker_3d.cpp:124-126), `rounds` x linearize
  const FramePointPointerVector& points = s->world->currentFrame()->points();
  const int n = (int)points.size();
  const double t0 = srrg_core::getTime();
  if (n && !s->gpu) {
    UV* a = s->uv;
    const double max_depth = s->parameters->stereo_framepoint_generator_parameters->maximum_reliable_depth_meters;
    a->_frame_previous = nullptr;
    a->_frame_current = s->world->currentFrame();
    a->_number_of_measurements = n;
    a->_errors.resize(n);
    a->_inliers.resize(n);
    a->_information_matrix_vector.resize(n);
    a->_weights_translation.resize(n);
    a->_moving.resize(n);
    a->_fixed.resize(n);
    for (int u = 0; u < n; ++u) {                                   // stereouv_aligner.cpp:26-64
      const FramePoint* p = points[u];
      a->_information_matrix_vector[u].setIdentity();
      a->_fixed[u] = Vector4(p->imageCoordinatesLeft().x(), p->imageCoordinatesLeft().y(), p->imageCoordinatesRight().x(),
                             p->imageCoordinatesRight().y());
      a->_moving[u] = p->cameraCoordinatesLeft();
      a->_weights_translation[u] = std::min(max_depth / p->depthMeters(), 1.0);
    }
    a->_camera_calibration_matrix = s->camera_left->cameraMatrix();
    a->_offset_camera_right = s->camera_right->baselineHomogeneous();
    a->_number_of_rows_image = s->rows;
    a->_number_of_cols_image = s->cols;
    a->_minimum_reliable_depth_meters = s->parameters->stereo_framepoint_generator_parameters->minimum_depth_meters;
    a->_previous_to_current = to_transform(T);
    for (int r = 0; r < rounds; ++r) a->linearize(false);
  }
  if (seconds_pose_optimization) *seconds_pose_optimization += srrg_core::getTime() - t0;
  return n;
  REF_CATCH(-1)
}

// ---- tracker -------------------------------------------------------------------------------------------------------
extern "C" int ref_tracker_process(ref_session* s, const uint8_t* left, const uint8_t* right, int stride) {
  REF_TRY
  // SLAMAssembly::process (slam_assembly.cpp:554-575, stereo branch, without relocalisation / map optimisation)
  s->tracker->setIntensityImageLeft(copy_image(left, s->rows, s->cols, stride));
  s->tracker->setImageSecondary(copy_image(right, s->rows, s->cols, stride));
  if (s->uv) s->uv->rounds = 0;
  s->tracker->compute();
  return (int)s->world->currentFrame()->points().size();
  REF_CATCH(-1)
}
extern "C" int ref_tracker_pose(ref_session* s, double robot_to_world[12]) {
  REF_TRY
  from_transform(s->world->currentFrame()->robotToWorld(), robot_to_world);
  return 0;
  REF_CATCH(-1)
}
extern "C" int ref_tracker_status(ref_session* s) { return s->tracker->_status == Frame::Tracking ? 1 : 0; }
extern "C" int ref_tracker_counts(ref_session* s, int* tracked_points, int* landmarks, int* frame_points) {
  REF_TRY
  if (tracked_points) *tracked_points = (int)s->tracker->_number_of_tracked_points;
  if (landmarks) *landmarks = (int)s->tracker->_number_of_active_landmarks;
  if (frame_points) *frame_points = (int)s->world->currentFrame()->points().size();
  return 0;
  REF_CATCH(-1)
}
extern "C" double ref_tracker_seconds(ref_session* s, int which) {
  switch (which) {
    case 0: return s->tracker->getTimeConsumptionSeconds_tracking();
    case 1: return s->tracker->getTimeConsumptionSeconds_track_creation();
    case 2: return s->tracker->getTimeConsumptionSeconds_pose_optimization();
    case 3: return s->tracker->getTimeConsumptionSeconds_landmark_optimization();
    default: return s->tracker->getTimeConsumptionSeconds_point_recovery();
  }
}
extern "C" int ref_write_trajectory(ref_session* s, int format, const char* filename) {
  REF_TRY
  if (format == 0) s->world->writeTrajectoryKITTI(filename);
  else s->world->writeTrajectoryTUM(filename);
  return 0;
  REF_CATCH(-1)
}

// ---- Landmark::update -------------------------------------------------------------------------------------------------
extern "C" int ref_landmark_run(ref_session* s, int n, const int32_t* frame_index, const double* camera_coordinates,
                                int n_frames, const double* robot_to_world, double world[3], uint32_t* number_of_updates) {
  REF_TRY
  std::vector<Frame*> frames(n_frames);
  for (int f = 0; f < n_frames; ++f) {
    frames[f] = s->world->createFrame();
    frames[f]->setCameraLeft(s->camera_left);
    frames[f]->setCameraRight(s->camera_right);
    frames[f]->setRobotToWorld(to_transform(robot_to_world + 12 * f));
  }
  IntensityFeature* feature = new IntensityFeature(cv::KeyPoint(cv::Point2f(0, 0), 7), cv::Mat(1, 32, CV_8UC1, cv::Scalar(0)), 0);
  s->dummy_features.push_back(feature);
  Landmark* landmark = nullptr;
  FramePoint* previous = nullptr;
  for (int m = 0; m < n; ++m) {
    Frame* f = frames[frame_index[m]];
    const PointCoordinates c(camera_coordinates[3 * m], camera_coordinates[3 * m + 1], camera_coordinates[3 * m + 2]);
    FramePoint* p = f->createFramepoint(feature, feature, 0, c, previous);     // frame.cpp:61-84: robot / world coordinates
    f->points().push_back(p);
    // a landmark is born from a point with a previous one (minimum_track_length_for_landmark_creation = 2,
    // pose_tracker_3d.cpp:492-503): its constructor walks the track backwards (landmark.cpp:20-33)
    if (m == 0) { previous = p; continue; }
    if (!landmark) landmark = s->world->createLandmark(p);                      // pose_tracker_3d.cpp:503
    else landmark->update(p);                                                   // pose_tracker_3d.cpp:510
    previous = p;
  }
  for (int k = 0; k < 3; ++k) world[k] = landmark->coordinates()(k);
  if (number_of_updates) *number_of_updates = landmark->numberOfUpdates();
  return (int)landmark->numberOfUpdates();
  REF_CATCH(-1)
}
