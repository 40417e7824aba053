/*
 * oracle/ref/ref_harness.h -- C entry points of oracle/_ref/libvslam_ref.so: the reference's OWN, UNMODIFIED translation
 * units (compiled from where they lie under /root/reference by oracle/Makefile, target `_ref`, against the functional
 * stand-ins for Eigen / OpenCV / srrg_core / yaml-cpp in oracle/shims/) behind plain pointers and sizes, so that Python
 * tests can run the reference's classes on the same inputs as the tier-A restatement and the CUDA path.
 *
 * TEST INFRASTRUCTURE ONLY: tests/ and bench.py's CPU-baseline legs load it; the product never does.
 *
 * What runs behind each call is the reference's code, not a restatement:
 *   ref_fpg_*      StereoFramePointGenerator::{configure, initialize, track, recoverPoints, compute}
 *                  (src/framepoint_generation/stereo_framepoint_generator.cpp, base_framepoint_generator.cpp,
 *                  intensity_feature_matcher.cpp), Frame / FramePoint (src/types/frame.cpp, frame_point.cpp)
 *   ref_aligner_*  StereoUVAligner / UVDAligner::{initialize, linearize, oneRound, converge}
 *                  (src/aligners/stereouv_aligner.cpp, uvd_aligner.cpp)
 *   ref_tracker_*  PoseTracker3D::compute (src/position_tracking/pose_tracker_3d.cpp) over WorldMap / Landmark
 *                  (src/types/world_map.cpp, landmark.cpp, local_map.cpp), wired like SLAMAssembly::_createStereoTracker
 *                  (src/system/slam_assembly.cpp:48-76)
 *   ref_landmark_* Landmark::update (src/types/landmark.cpp:66-167)
 *   ref_open       ParameterCollection::parseFromFile (src/types/parameters.cpp:272-440), the reference's own parser
 * Third-party arithmetic is the stand-ins' (oracle/shims): FAST / ORB = tier A or an installed backend (cv2), Eigen
 * products in one documented order, FullPivLU with complete pivoting, srrg_core::v2t / skew from their definitions.
 */
#ifndef VSLAM_REF_HARNESS_H
#define VSLAM_REF_HARNESS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ref_session ref_session;

/* effective values of the hot path's parameters after the reference's own parsing (src/types/parameters.h:66-95,161-238)
 * -- read with ref_get_parameters, overridden with ref_set_parameters BEFORE ref_configure */
typedef struct {
  char detector_type[32], descriptor_type[32];
  double target_number_of_keypoints_tolerance;
  double detector_threshold_minimum, detector_threshold_maximum, detector_threshold_maximum_change;
  uint32_t number_of_detectors_vertical, number_of_detectors_horizontal;
  int32_t minimum_projection_tracking_distance_pixels, maximum_projection_tracking_distance_pixels;
  double minimum_descriptor_distance_tracking, maximum_descriptor_distance_tracking;
  double maximum_reliable_depth_meters, maximum_depth_meters, minimum_depth_meters;
  int32_t enable_keypoint_binning;
  uint32_t bin_size_pixels;
  double maximum_matching_distance_triangulation, minimum_disparity_pixels;
  int32_t maximum_epipolar_search_offset_pixels;
  int32_t use_matches;                       /* the dead FLANN / findHomography block (stereo :168-273) */
  /* tracker_parameters->aligner (AlignerParameters) */
  double error_delta_for_convergence, maximum_error_kernel, damping;
  uint32_t maximum_number_of_iterations, minimum_number_of_inliers;
  double minimum_inlier_ratio;
  int32_t enable_inverse_depth_as_information;
  /* PoseTracker3DParameters */
  uint32_t minimum_track_length_for_landmark_creation, minimum_number_of_landmarks_to_track;
  double tunnel_vision_ratio, good_tracking_ratio;
  uint32_t maximum_number_of_landmark_recoveries;
  int32_t enable_landmark_recovery;
  int32_t motion_model;                      /* 0 NONE, 1 CONSTANT_VELOCITY, 2 CAMERA_ODOMETRY */
  double minimum_delta_angular_for_movement, minimum_delta_translational_for_movement;
  /* LandmarkParameters */
  double maximum_error_squared_meters;
} ref_parameters;

/* one FramePoint of frame->points() */
typedef struct {
  float xl, yl, xr, yr;                 /* keypointLeft().pt, keypointRight().pt */
  int32_t row, col;
  int32_t epipolar_offset;
  int32_t index_previous;               /* position of previous() in the previous frame's points(), -1 if none */
  double disparity, distance;           /* disparityPixels(), descriptorDistanceTriangulation() */
  double cam[3], robot[3], world[3];
  float projection_left[2], projection_right[2], projection_right_corrected[2];
  int32_t has_landmark;
  uint32_t track_length;
  double landmark_world[3];             /* landmark()->coordinates() when has_landmark */
  uint32_t landmark_updates;            /* landmark()->numberOfUpdates() */
  int32_t reserved;
  uint8_t desc_left[32], desc_right[32];
} ref_point;

const char* ref_last_error(void);

/* yaml may be NULL (struct defaults).  K row-major 3x3; bx = P_right(0,3) = -fx * baseline (< 0). */
ref_session* ref_open(const char* yaml, int rows, int cols, const double K[9], double bx);
void ref_close(ref_session*);
int ref_get_parameters(ref_session*, ref_parameters* out);
int ref_set_parameters(ref_session*, const ref_parameters* in);
/* builds generator, aligners, world map and tracker exactly as SLAMAssembly::_createStereoTracker does */
int ref_configure(ref_session*);
/* the same wiring with adapters/ GpuStereoFramePointGenerator + GpuStereoUVAligner (libvslam_b200.so) in place of the CPU
 * classes -- only in oracle/_ref/libvslam_ref_gpu.so (make _ref_gpu); calls that read CPU internals then fail */
int ref_configure_gpu(ref_session*);
int ref_has_gpu_adapters(void);

/* ---- generator, stage by stage ------------------------------------------------------------------------------ */
/* WorldMap::createFrame + setStatus(status: 0 Localizing, 1 Tracking) + StereoFramePointGenerator::initialize */
int ref_fpg_initialize(ref_session*, const uint8_t* left, const uint8_t* right, int stride, int status);
/* initialize(frame, false) on the current frame (the tracker's retry, pose_tracker_3d.cpp:319) */
int ref_fpg_reinitialize(ref_session*);
int ref_fpg_features(ref_session*, int side, float* xyr, uint8_t* desc, int capacity);     /* frame->keypoints/descriptors */
int ref_fpg_remaining(ref_session*, int side, float* xy, int capacity);                   /* matcher.feature_vector */
int ref_fpg_thresholds(ref_session*, double* out, int capacity);                          /* detectors' thresholds */
double ref_fpg_triangulation_distance(ref_session*);
int ref_fpg_target_number_of_keypoints(ref_session*);
int ref_fpg_set_tracking(ref_session*, int projection_tracking_distance_pixels, double maximum_descriptor_distance_tracking);
/* track(current, previous, T, lost, by_appearance); lost[] = positions in the previous frame's points() */
int ref_fpg_track(ref_session*, const double previous_to_current[12], int track_by_appearance, int32_t* lost, int* n_lost,
                  int* number_of_tracked_landmarks, double* average_descriptor_distance);
int ref_fpg_recover(ref_session*);              /* recoverPoints(current, lost of the last track) -> number of points added */
int ref_fpg_compute(ref_session*);              /* -> frame->points().size() */
int ref_frame_points(ref_session*, int which /*0 current, 1 previous*/, ref_point* out, int capacity);
int ref_frame_set_pose(ref_session*, const double robot_to_world[12]);      /* Frame::setRobotToWorld on the current frame */
/* promote the current frame's points to landmarks with `updates` updates each (WorldMap::createLandmark +
 * Landmark::update of the reference), for the landmark branches of track() / recoverPoints() / initialize() */
int ref_frame_make_landmarks(ref_session*, int every_nth);
double ref_fpg_seconds(ref_session*, int which /*0 detection, 1 description, 2 triangulation*/);

/* ---- aligners on explicit correspondences (BASELINE configs[3]) ---------------------------------------------- */
/* kind 0 = StereoUVAligner, 1 = UVDAligner.  Arrays as oracle/c/vslam_oracle.h's orc_aligner_problem. */
int ref_aligner_load(ref_session*, int kind, int n, const double* moving, const double* fixed, const double* omega,
                     const double* wt, const double baseline[3], double min_depth);
int ref_aligner_set_pose(ref_session*, int kind, const double T[12]);
int ref_aligner_linearize(ref_session*, int kind, int ignore_outliers);
int ref_aligner_one_round(ref_session*, int kind, int ignore_outliers);
int ref_aligner_converge(ref_session*, int kind);       /* returns hasSystemConverged */
/* state after any of the above; NULL pointers are skipped */
int ref_aligner_state(ref_session*, int kind, double H[36], double b[6], double* total_error, int* inliers, int* outliers,
                      double T[12], double* errors, uint8_t* inlier_flags, double information[36]);
/* the aligner's own initialize(previous, current, T0) on the session's frames; then dump what it packed */
int ref_aligner_initialize_frames(ref_session*, int kind, const double T0[12], int enable_inverse_depth_as_information);
int ref_aligner_packed(ref_session*, int kind, double* moving, double* fixed, double* omega, double* wt);
int ref_aligner_count(ref_session*, int kind);
int ref_aligner_rounds(ref_session*, int kind);         /* oneRound calls since the last load / initialize */

/* ---- bench.py's batched first-frame workload (BASELINE configs[2]) on the reference's classes ---------------------- */
int ref_reset(ref_session*);    /* a fresh sequence: world cleared, detector thresholds back at their minimum */
/* reset -> initialize(Localizing) -> compute -> StereoUVAligner packed from the frame's points, `rounds` x linearize at T */
int ref_first_frame(ref_session*, const uint8_t* left, const uint8_t* right, int stride, int rounds, const double T[12],
                    double* seconds_pose_optimization);

/* ---- the whole tracker (BASELINE configs[0]: executables/app's per-frame call) ------------------------------- */
int ref_tracker_process(ref_session*, const uint8_t* left, const uint8_t* right, int stride);
int ref_tracker_pose(ref_session*, double robot_to_world[12]);
int ref_tracker_status(ref_session*);                    /* 0 Localizing, 1 Tracking */
int ref_tracker_counts(ref_session*, int* tracked_points, int* landmarks, int* frame_points);
double ref_tracker_seconds(ref_session*, int which /*0 tracking, 1 track_creation, 2 pose_optimization, 3 landmark_optimization, 4 point_recovery*/);
int ref_write_trajectory(ref_session*, int format /*0 KITTI, 1 TUM*/, const char* filename);

/* ---- Landmark::update on an explicit measurement history ------------------------------------------------------ */
/* measurements: n x (frame index, camera coordinates[3]); poses: robot_to_world per frame (camera == robot).  Runs
 * createLandmark on the first point and update() for each later one; returns the number of updates; world[3] out. */
int ref_landmark_run(ref_session*, int n, const int32_t* frame_index, const double* camera_coordinates,
                     int n_frames, const double* robot_to_world, double world[3], uint32_t* number_of_updates);

#ifdef __cplusplus
}
#endif
#endif
