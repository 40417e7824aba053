/*
 * vslam_b200.h -- C ABI of libvslam_b200.so: the B200-native (sm_100a) replacement of the per-frame hot
 * path of Ssellu/vslam-pose-estimation-framework (ProSLAM fork):
 *
 *   stereo framepoint generation   src/framepoint_generation/{base,stereo}_framepoint_generator.cpp
 *   projective aligner linearise   src/aligners/{stereouv,uvd}_aligner.cpp
 *
 * Every entry point cites the reference interface it replaces (file:line relative to the reference root).
 * The reference has no FFI: its boundary is two abstract C++ classes injected by raw pointer
 * (src/position_tracking/pose_tracker_3d.h:36-37, src/system/slam_assembly.cpp:62-76).  The C++14 adapters
 * in adapters/ derive from those classes and call ONLY the functions below; INTEGRATION.md shows the wiring.
 *
 * Conventions: opaque handles; plain pointers and sizes; caller-owned HOST buffers unless a name says
 * "device"; int status (0 ok, <0 error, message via vslam_last_error()); no exceptions cross the boundary;
 * one handle = one CUDA device + one stream, not thread-safe per handle; there is NO CPU fallback --
 * every call that computes fails with VSLAM_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef VSLAM_B200_H
#define VSLAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSLAM_OK 0
#define VSLAM_ERR_INVALID_ARGUMENT (-1)
#define VSLAM_ERR_CUDA (-2)
#define VSLAM_ERR_CAPACITY (-3) /* more keypoints / points than the handle was created for */
#define VSLAM_ERR_STATE (-4)    /* call order violated (e.g. compute before initialize) */

#define VSLAM_MAX_DETECTOR_REGIONS 64
#define VSLAM_DESCRIPTOR_BYTES 32 /* SRRG_PROSLAM_DESCRIPTOR_SIZE_BITS = 256, CMakeLists.txt:32 */

/* message of the last failing call on the calling thread */
const char* vslam_last_error(void);
/* "vslam_b200 <version> sm_100a" ; number of usable CUDA devices (0 when none: nothing will compute) */
const char* vslam_version(void);
int vslam_device_count(void);

/* pinned host memory for the caller-owned buffers of the batched calls (optional; any host pointer works,
 * pinned ones make the copies asynchronous) */
int vslam_host_alloc(void** ptr, size_t bytes);
int vslam_host_free(void* ptr);

/* the normal equations BaseAligner::linearize leaves behind (AlignerWorkspace<6,D>, base_aligner.h:74-106) */
typedef struct {
  double H[36];            /* _H, row-major */
  double b[6];             /* _b */
  double total_error;      /* _total_error */
  int32_t number_of_inliers;
  int32_t number_of_outliers;
} vslam_linear_system;

/* ===================================================================================================
 * Stereo framepoint generation
 * =================================================================================================*/

typedef struct vslam_fpg vslam_fpg;

/* The reference's parameter structs, unchanged in meaning:
 * BaseFramePointGeneratorParameters / StereoFramePointGeneratorParameters (src/types/parameters.h:161-238),
 * plus what configure() reads from the cameras (base_framepoint_generator.cpp:169-173,
 * stereo_framepoint_generator.cpp:26-38). */
typedef struct {
  int32_t rows, cols;                                /* _camera_left->numberOfImageRows/Cols */
  double target_number_of_keypoints_tolerance;
  int32_t detector_threshold_minimum;
  int32_t detector_threshold_maximum;
  double detector_threshold_maximum_change;
  int32_t number_of_detectors_vertical;
  int32_t number_of_detectors_horizontal;
  int32_t enable_keypoint_binning;
  int32_t bin_size_pixels;
  double maximum_matching_distance_triangulation;
  double minimum_disparity_pixels;
  int32_t maximum_epipolar_search_offset_pixels;
  double fx, fy, cx, cy;                             /* _camera_right->cameraMatrix() */
  double bx;                                         /* _camera_right->baselineHomogeneous()(0), < 0 */
  /* capacities of the device-resident state (not reference parameters) */
  int32_t max_keypoints_per_image;                   /* descriptor-valid keypoints per image; 0 -> default */
  int32_t max_batch;                                 /* stereo pairs per batched call; 0 -> 1 */
  /* descriptor extractor, base_framepoint_generator.cpp:184-224.  VSLAM_DESCRIPTOR_ORB (the default, 0): cv::ORB::create()
   * -- what "ORB", "ORB-256", "BRIEF-256" and, without opencv_contrib, "BRIEF" resolve to (:187-195, :219-224).
   * VSLAM_DESCRIPTOR_BRIEF: cv::xfeatures2d::BriefDescriptorExtractor::create(32) (:186, "BRIEF" with opencv_contrib);
   * brief_tests then points to its 256 x 4 test table (y0, x0, y1, x1 of `SMOOTHED(y0, x0) < SMOOTHED(y1, x1)`, in the
   * order of opencv_contrib's generated_32.i, offsets within +-24), copied at creation. */
  int32_t descriptor_type;
  int32_t reserved;
  const int8_t* brief_tests;
} vslam_fpg_config;

#define VSLAM_DESCRIPTOR_ORB 0
#define VSLAM_DESCRIPTOR_BRIEF 1

/* cv::KeyPoint as the reference sees it after computeDescriptors (size 7, angle -1, octave 0, class_id -1
 * are implied: cv::FastFeatureDetector defaults) */
typedef struct {
  float x, y;      /* pt */
  float response;  /* FAST corner score */
} vslam_keypoint;

/* what StereoFramePointGenerator::compute hands to Frame::createFramepoint
 * (stereo_framepoint_generator.cpp:364-368; src/types/frame_point.cpp:8-24) */
typedef struct {
  int32_t index_left, index_right; /* rows of keypointsLeft()/descriptorsLeft() resp. Right (reference order) */
  float xl, yl, xr, yr;            /* keypoint.pt of both features */
  int32_t distance;                /* descriptor_distance_triangulation (Hamming) */
  int32_t epipolar_offset;         /* setEpipolarOffset */
  double camera[3];                /* getPointInLeftCamera */
} vslam_framepoint;

/* a point already in frame->points() when compute() starts (stereo_framepoint_generator.cpp:147-155) */
typedef struct {
  int32_t row, col;       /* FramePoint::row / col */
  int32_t has_previous;   /* previous() != nullptr */
  int32_t reserved;
  double disparity;       /* disparityPixels() */
  double distance;        /* descriptorDistanceTriangulation() */
} vslam_tracked_point;

/* StereoFramePointGenerator ctor + configure(): stereo_framepoint_generator.cpp:7-60,
 * base_framepoint_generator.cpp:165-329.  Fails like the reference on a non-positive baseline (:29-34). */
int vslam_fpg_create(const vslam_fpg_config* config, int device, vslam_fpg** out);
int vslam_fpg_destroy(vslam_fpg* h);

/* derived configuration (base_framepoint_generator.cpp:293-312): regions as x,y,w,h quadruples */
int vslam_fpg_info(const vslam_fpg* h, int32_t* n_regions, int32_t* regions_xywh, int32_t* rows_bin,
                   int32_t* cols_bin, int32_t* target_number_of_keypoints);

/* detector thresholds, one per region, row-major (FastDetector::getThreshold/setThreshold, :16-22) */
int vslam_fpg_get_thresholds(const vslam_fpg* h, double* thresholds);
int vslam_fpg_set_thresholds(vslam_fpg* h, const double* thresholds);

/* StereoFramePointGenerator::initialize(frame, extract_features=true), :73-133:
 * detectKeypoints(L), detectKeypoints(R), adjustDetectorThresholds(), computeDescriptors(L/R), triangulation
 * distance for frame->status() (localizing != 0 <=> Frame::Localizing), setFeatures(L/R).
 * left/right: rows x cols uint8, `stride` bytes between rows.  Results stay on the device;
 * n_left/n_right = keypointsLeft()/Right().size() after the call. */
int vslam_fpg_initialize(vslam_fpg* h, const uint8_t* left, const uint8_t* right, size_t stride, int localizing,
                         int32_t* n_left, int32_t* n_right);

/* frame->keypointsLeft()/descriptorsLeft() (side 0) or ...Right() (side 1) after initialize, in the
 * reference's order (detector region by region, row-major inside a region).  descriptors: n x 32 bytes. */
/* how many vslam_fpg_initialize calls ran their device side as ONE CUDA-graph launch (threshold upload, repitch, FAST,
 * compact, blur, descriptors, status download captured once per handle and profiling mode; used when the images are
 * uploaded linearly, i.e. unless their rows are very widely strided).  Identical kernels and results either way. */
int64_t vslam_fpg_graph_launch_count(const vslam_fpg* h);

int vslam_fpg_get_features(vslam_fpg* h, int side, vslam_keypoint* keypoints, uint8_t* descriptors,
                           int32_t capacity, int32_t* n);

/* raw FAST keypoint count per detector region of the last initialize (what drives the threshold
 * controller, base_framepoint_generator.cpp:382), and _current_maximum_descriptor_distance_triangulation */
int vslam_fpg_get_detection_stats(vslam_fpg* h, int32_t* counts_left, int32_t* counts_right,
                                  double* matching_distance);

/* What track() leaves behind (stereo_framepoint_generator.cpp:646-651, 671-672: matched features are pruned from
 * _feature_matcher_left/right before compute() scans the rest).  `remaining` = keypoints of the features still in
 * the matcher's feature_vector for `side` (0 left, 1 right); every other feature of the frame is excluded from the
 * next compute().  Without this call every feature of initialize() takes part (first frame / Localizing restart). */
int vslam_fpg_set_remaining_features(vslam_fpg* h, int side, const vslam_keypoint* remaining, int32_t n);

/* StereoFramePointGenerator::initialize(frame, extract_features = false) (stereo_framepoint_generator.cpp:73-84,
 * 126-133): no new detection; both feature matchers are set up again from the frame's keypoints and descriptors, i.e.
 * every feature of the last vslam_fpg_initialize is available again and what a track() attempt pruned is forgotten.
 * PoseTracker3D retries a failed track() this way (pose_tracker_3d.cpp:320, 402). */
int vslam_fpg_reset_features(vslam_fpg* h);

/* What track() and recoverPoints() read of one FramePoint of the previous frame
 * (stereo_framepoint_generator.cpp:494-606, 702-835; src/types/frame_point.h) */
typedef struct {
  double camera_left[3];        /* cameraCoordinatesLeft()                                   (track) */
  double world[3];              /* landmark()->coordinates(), world frame                    (recoverPoints) */
  uint8_t descriptor_left[32];  /* descriptorLeft() */
  uint8_t descriptor_right[32]; /* descriptorRight() */
  int32_t epipolar_offset;      /* epipolarOffset() */
  int32_t has_landmark;         /* landmark() != nullptr */
  float keypoint_size;          /* keypointLeft().size (7 for cv::FAST keypoints)            (recoverPoints) */
  int32_t reserved;
} vslam_previous_point;

/* one tracked and triangulated point: the arguments of Frame::createFramepoint(feature_left, feature_right,
 * descriptor_distance_best, getPointInLeftCamera(...), point_previous) and the setters behind it (:623-643) */
typedef struct {
  int32_t index_previous;          /* position of point_previous in frame_previous->points() */
  int32_t index_left, index_right; /* rows of keypointsLeft()/descriptorsLeft() resp. Right (reference order) */
  float xl, yl, xr, yr;            /* keypoint.pt of both features */
  int32_t distance;                /* descriptor_distance_best (of the right search, :584-590) */
  int32_t epipolar_offset;         /* setEpipolarOffset(feature_right->row - feature_left->row) */
  float projection_left[2];        /* setProjectionEstimateLeft */
  float projection_right[2];       /* setProjectionEstimateRight */
  float projection_right_corrected[2]; /* setProjectionEstimateRightCorrected */
  int32_t reserved;
  double camera[3];                /* getPointInLeftCamera */
} vslam_track;

/* one recovered point (:840-858): keypoints at the rounded projections, the two new descriptors, their distance */
typedef struct {
  int32_t index_lost;              /* position in the lost-point array */
  int32_t distance;                /* descriptor_distance_triangulation */
  float xl, yl, xr, yr;
  double camera[3];
  uint8_t descriptor_left[32], descriptor_right[32];
} vslam_recovered_point;

/* StereoFramePointGenerator::track(frame, frame_previous, camera_left_previous_in_current, lost_points,
 * track_by_appearance), :464-681, after initialize() of the current frame.  projection_tracking_distance_pixels and
 * maximum_descriptor_distance_tracking are what setProjectionTrackingDistancePixels /
 * setMaximumDescriptorDistanceTracking (base_framepoint_generator.h:164-165) set before the call.
 * tracks: frame->points() after the call, in order.  lost: positions (in `previous`) of lost_points_, in order.
 * average_descriptor_distance: what setAverageDescriptorDistanceTracking receives (NaN without tracks).
 * The matched features (and the right features in the parallax ranges) are pruned on the device exactly like
 * :671-672, so the next compute() scans only the rest; no vslam_fpg_set_remaining_features call is needed. */
int vslam_fpg_track(vslam_fpg* h, const vslam_previous_point* previous, int32_t n_previous,
                    const double previous_to_current[12], int track_by_appearance,
                    int32_t projection_tracking_distance_pixels, double maximum_descriptor_distance_tracking,
                    vslam_track* tracks, int32_t capacity, int32_t* n_tracks, int32_t* lost, int32_t* n_lost,
                    int32_t* n_tracked_landmarks, double* average_descriptor_distance);

/* PoseTracker3D::_prunePoints (src/position_tracking/pose_tracker_3d.cpp:437-472) for the device-resident tracks: after
 * the aligner has converged on the tracks of the last vslam_fpg_track (correspondence k = track k), the tracks it rejects
 * are dropped from the records that pre-load the bins of vslam_fpg_compute(VSLAM_TRACKED_FROM_LAST_TRACK) -- on the
 * device, in order, without the tracked points travelling host -> device again.  Rule: average error < kernel ? inliers
 * only : errors != -1 and < 100 kernel.  n_kept / kept[n_tracks] (either may be NULL) tell the host which of its own
 * track records to keep (frame->points().resize, :471).  Declared here, after both handle types. */
typedef struct vslam_aligner vslam_aligner;
int vslam_fpg_prune_tracks(vslam_fpg* h, vslam_aligner* aligner, double maximum_error_kernel, int32_t* n_kept,
                           uint8_t* kept);

/* StereoFramePointGenerator::recoverPoints(frame, lost_points), :683-869, with the ORB extractor.  recovered: the
 * points appended to frame->points(), in order.  minimum/maximum_depth_meters: parameters.h:197-198. */
int vslam_fpg_recover_points(vslam_fpg* h, const vslam_previous_point* lost, int32_t n_lost,
                             const double world_to_camera_left[12], double minimum_depth_meters,
                             double maximum_depth_meters, double maximum_descriptor_distance_tracking,
                             vslam_recovered_point* recovered, int32_t capacity, int32_t* n_recovered);

/* AlignerParameters (src/types/parameters.h:66-95) + BaseAligner thresholds (base_aligner.h:62-63) */
typedef struct {
  double error_delta_for_convergence;
  double maximum_error_kernel;
  double damping;
  int32_t maximum_number_of_iterations;
  int32_t minimum_number_of_inliers;
} vslam_aligner_parameters;

/* ---- One tracked frame as ONE device pass ------------------------------------------------------------------------
 * PoseTracker3D::compute's per-frame order (src/position_tracking/pose_tracker_3d.cpp): :80 initialize, :239 track
 * against points() of the previous frame, :124-126 / :355-357 StereoUVAligner::initialize + converge over the tracks,
 * :437-472 _prunePoints, :210 compute.  The stepwise calls above return to the host between the stages because the
 * reference's tracker owns that order; a host that owns its loop calls this instead: points() of the previous frame
 * stay on the device (camera coordinates + both descriptors, 128 B each), the aligner's correspondences are packed on
 * the device (stereouv_aligner.cpp:26-64, the branch without a landmark estimate: moving =
 * previous->cameraCoordinatesLeft(), information = I), converge() runs as one thread-block cluster, the prune rule and
 * the bin pre-load never leave the device, and the whole frame is one CUDA-graph launch (kernels that do not depend on
 * each other run as parallel branches of the graph) and ONE wait: the call returns when the frame's last kernel has
 * echoed the frame number into the pinned result block, which the host polls.
 * Results are bit-identical to the stepwise calls in the same order (tests/test_gpu_frame_step.py).
 * Limits: keypoint binning enabled; at most vslam_fpg_frame_step_capacity() points per frame (4096 on a B200: one
 * 16-CTA cluster, one correspondence per thread) -- VSLAM_ERR_CAPACITY beyond, the stepwise calls have no such limit.
 * index_left / index_right of the returned records count in the device's (row, col) order, which is the reference's
 * keypoint order for one detector region; the descriptors travel inside frame_points. */
typedef struct {
  int32_t track_by_appearance;                       /* track(..., track_by_appearance_) */
  int32_t projection_tracking_distance_pixels;       /* base_framepoint_generator.h:164 */
  double maximum_descriptor_distance_tracking;       /* :165 */
  vslam_aligner_parameters aligner;                  /* AlignerParameters of the pose optimiser */
  int32_t enable_inverse_depth_as_information;       /* parameters.h:88 */
  int32_t minimum_track_length_for_landmark_creation;/* parameters.h:259: has_landmark of the frame's points (track length >= this) */
  double maximum_reliable_depth_meters;              /* slam_assembly.cpp:70 */
  double minimum_reliable_depth_meters;              /* slam_assembly.cpp:69 */
  int32_t publish_frame_points;                      /* != 0: frame_points below is filled (128 B per point) */
  int32_t reserved;
} vslam_frame_step_parameters;

typedef struct {
  int32_t n_left, n_right;                 /* descriptor-valid keypoints of the frame */
  int32_t n_previous;                      /* points of the previous frame the frame was tracked against */
  int32_t n_tracked, n_lost;               /* track(): tracks before the prune, lost points */
  int32_t n_tracked_landmarks;             /* _number_of_tracked_landmarks */
  int32_t n_tracks;                        /* tracks that survive _prunePoints */
  int32_t n_new_points, n_matches;         /* compute(): framepoints appended, matches before binning */
  int32_t aligner_rounds, aligner_converged, aligner_inliers, aligner_outliers;
  int32_t inliers_only;                    /* the branch of pose_tracker_3d.cpp:441 */
  double average_descriptor_distance;      /* setAverageDescriptorDistanceTracking (NaN without tracks) */
  double aligner_total_error;
  double previous_to_current[12];          /* the optimised motion (the prior when there was nothing to align) */
  double information[36];                  /* _information_matrix (damped H of the last round) */
  /* pinned memory owned by the handle, valid until the next call on it */
  const vslam_track* tracks;               /* [n_tracks] surviving tracks, in order */
  const uint8_t* kept;                     /* [n_tracked] 1 = track k survived */
  const double* errors;                    /* [n_tracked] _errors of the last round */
  const uint8_t* inliers;                  /* [n_tracked] _inliers */
  const int32_t* lost;                     /* [n_lost] positions in the previous frame's points() */
  const vslam_framepoint* points;          /* [n_new_points] */
  const vslam_previous_point* frame_points;/* [n_tracks + n_new_points] points() of this frame (publish_frame_points) */
} vslam_frame_step_result;

int32_t vslam_fpg_frame_step_capacity(vslam_fpg* h);
/* a new sequence: no previous points */
int vslam_fpg_frame_step_reset(vslam_fpg* h);
/* points() of the previous frame from the host (e.g. after the host re-estimated them from its landmarks) */
int vslam_fpg_frame_step_set_previous(vslam_fpg* h, const vslam_previous_point* previous, int32_t n_previous);
int vslam_fpg_frame_step(vslam_fpg* h, const uint8_t* left, const uint8_t* right, size_t stride_bytes, int localizing,
                         const double previous_to_current_prior[12], const vslam_frame_step_parameters* parameters,
                         vslam_frame_step_result* result);
/* Landmark estimates for the points() the device holds (the frame that just returned), read by the NEXT frame's pose
 * optimisation as StereoUVAligner::initialize does (stereouv_aligner.cpp:43-51): a point whose entry has
 * information_scale != 0 is moved as `camera` (previous->cameraCoordinatesLeftLandmark()) with its information scaled by
 * information_scale (the caller's 1 + log(landmark->numberOfUpdates())); 0 = no landmark, the point's own camera
 * coordinates and identity information.  One entry per point, in the order of frame_points (n_tracks survivors, then the
 * new points); entries are forgotten when the next frame replaces points().  Landmark refinement itself stays with the
 * host or with the device-resident landmark map below. */
typedef struct {
  double camera[3];
  double information_scale;
} vslam_landmark_estimate;
int vslam_fpg_frame_step_set_landmark_estimates(vslam_fpg* h, const vslam_landmark_estimate* estimates, int32_t n_points);
/* Upload of the NEXT frame's images while the current frame runs (a host that replays a sequence knows them: the reference's
 * playback, executables/app.cpp, reads them from disk).  Asynchronous on a copy stream of the handle; the images must stay
 * valid (and should be page-locked, vslam_host_alloc) until the vslam_fpg_frame_step that consumes them returns.  Staged
 * pairs are consumed oldest first by vslam_fpg_frame_step with left == right == NULL (stride_bytes is then ignored); at most
 * two pairs are staged (VSLAM_ERR_STATE beyond), and a call WITH images while a pair is staged is VSLAM_ERR_STATE too.
 * vslam_fpg_frame_step_reset drops staged pairs.  Results do not depend on how the images arrived. */
int vslam_fpg_frame_step_prefetch(vslam_fpg* h, const uint8_t* left, const uint8_t* right, size_t stride_bytes);

/* pass as n_tracked (tracked = NULL) to vslam_fpg_compute: the tracked points are those of the last
 * vslam_fpg_track, already resident on the device (no host round trip of the bin pre-load records) */
#define VSLAM_TRACKED_FROM_LAST_TRACK (-1)

/* StereoFramePointGenerator::compute(frame), :135-462 (without the dead use_matches block :168-273).
 * tracked: the points already in frame->points().  framepoints: the points compute() appends to
 * frame->points(), in order (bin winners without previous(), row-major over bins; every match in emission
 * order when binning is off).  n_matches (optional) = number_of_new_points before binning. */
int vslam_fpg_compute(vslam_fpg* h, const vslam_tracked_point* tracked, int32_t n_tracked,
                      vslam_framepoint* framepoints, int32_t capacity, int32_t* n_framepoints,
                      int32_t* n_matches);

/* all new stereo matches of the last compute in emission order (framepoints_new, :163,397) */
int vslam_fpg_get_matches(vslam_fpg* h, vslam_framepoint* matches, int32_t capacity, int32_t* n);

/* getTimeConsumptionSeconds_{keypoint_detection, descriptor_extraction, point_triangulation}
 * (base_framepoint_generator.h:232-233, stereo_framepoint_generator.h:81): accumulated DEVICE seconds,
 * measured with CUDA events when profiling is enabled. */
/* enabled > 0: on (batched calls then run serialised on one stream); 0: off; < 0: off and reset accumulators */
int vslam_fpg_set_profiling(vslam_fpg* h, int enabled);
/* accumulated device milliseconds and launch counts per kernel while profiling was on, in this order:
 * fast_nms, compact, blur, describe, match, select, linearize_pairs, track (search + resolve + emit),
 * frame aligner (StereoUVAligner initialize + converge + _prunePoints inside vslam_fpg_frame_step).  With profiling on,
 * vslam_fpg_frame_step runs its kernels as ONE chain (no parallel branches) with the timing events between them. */
#define VSLAM_FPG_KERNELS 9
int vslam_fpg_get_kernel_profile(vslam_fpg* h, double* milliseconds, int64_t* launches);
int vslam_fpg_get_time_consumption(vslam_fpg* h, double* keypoint_detection, double* descriptor_extraction,
                                   double* point_triangulation);

/* ---- batched form: n_pairs INDEPENDENT stereo pairs (BASELINE.json config 3) -------------------------
 * Every pair is processed exactly like initialize()+compute() on a fresh generator whose thresholds are
 * the handle's current ones (they are not updated by batched calls) and with no tracked points.
 * Images: pair i at left + i*pair_stride (rows x cols, row stride `stride`). */
int vslam_fpg_batch_upload(vslam_fpg* h, int32_t n_pairs, const uint8_t* left, const uint8_t* right,
                           size_t stride, size_t pair_stride);
/* all kernels of the path, device-resident inputs and outputs, asynchronous on the handle's stream */
int vslam_fpg_batch_run(vslam_fpg* h, int32_t n_pairs, int localizing);
/* framepoints of pair i at framepoints + i*capacity_per_pair; n_framepoints[i], n_matches[i] (optional),
 * n_left[i], n_right[i] (optional).  Synchronises the stream. */
int vslam_fpg_batch_download(vslam_fpg* h, int32_t n_pairs, vslam_framepoint* framepoints,
                             int32_t capacity_per_pair, int32_t* n_framepoints, int32_t* n_matches,
                             int32_t* n_left, int32_t* n_right);
/* upload + run + download with copy/compute overlap: the end-to-end call */
int vslam_fpg_batch_process(vslam_fpg* h, int32_t n_pairs, const uint8_t* left, const uint8_t* right,
                            size_t stride, size_t pair_stride, int localizing, vslam_framepoint* framepoints,
                            int32_t capacity_per_pair, int32_t* n_framepoints);
/* StereoUVAligner::initialize (stereouv_aligner.cpp:10-69) + `rounds` x linearize (:72-187) for every pair of the
 * last batched run, device-resident, one problem per pair: the pair's new framepoints are aligned against
 * themselves (what the tracker does when previous and current frame hold the same points and no landmarks):
 * _moving = cameraCoordinatesLeft, _fixed = (uL,vL,uR,vR), information = I4, w_t = min(max_depth/depth, 1) if
 * enable_inverse_depth_as_information else 1.  Asynchronous on the handle's stream. */
int vslam_fpg_batch_linearize(vslam_fpg* h, int32_t n_pairs, const double previous_to_current[12], int ignore_outliers,
                              double maximum_error_kernel, double minimum_reliable_depth,
                              double maximum_reliable_depth, int enable_inverse_depth_as_information, int32_t rounds);
/* systems[n_pairs]; optional per-point _errors / _inliers, pair i at + i*capacity_per_pair.  Synchronises. */
int vslam_fpg_batch_get_systems(vslam_fpg* h, int32_t n_pairs, vslam_linear_system* systems, double* errors,
                                uint8_t* inliers, int32_t capacity_per_pair);
/* per-pair features of the last batched run (reference order), for parity checks */
int vslam_fpg_batch_get_features(vslam_fpg* h, int32_t pair, int side, vslam_keypoint* keypoints,
                                 uint8_t* descriptors, int32_t capacity, int32_t* n);

/* the handle's cudaStream_t (as void*) and a stream synchronise, for callers that time with CUDA events */
void* vslam_fpg_stream(vslam_fpg* h);
int vslam_fpg_synchronize(vslam_fpg* h);
/* number of kernels the handle has launched so far */
int64_t vslam_fpg_launch_count(const vslam_fpg* h);

/* debug / parity taps on device intermediates of pair `pair`, image `side`:
 * FAST keypoint bitmask before the 31 px descriptor border filter (rows x ceil(cols/32) words, bit x%32 of
 * word x/32), and the 7x7 sigma 2 blurred image (rows x cols) cv::ORB::compute samples. */
int vslam_fpg_debug_keypoint_mask(vslam_fpg* h, int32_t pair, int side, uint32_t* words);
int vslam_fpg_debug_blurred(vslam_fpg* h, int32_t pair, int side, uint8_t* image);

/* BaseFramePointGenerator threshold controller for one region (base_framepoint_generator.cpp:377-415),
 * exposed so hosts that batch independent sequences can run it themselves */
double vslam_threshold_proposal(double threshold, int32_t n_keypoints, double target_per_detector,
                                double tolerance, double maximum_change, double threshold_minimum,
                                double threshold_maximum);

/* ===================================================================================================
 * Frame aligners (pose optimisation previous -> current)
 * =================================================================================================*/

/* (vslam_aligner is declared above, with vslam_fpg_prune_tracks) */

#define VSLAM_ALIGNER_STEREO_UV 0 /* src/aligners/stereouv_aligner.cpp */
#define VSLAM_ALIGNER_UVD 1       /* src/aligners/uvd_aligner.cpp */

/* (vslam_aligner_parameters is declared above, with vslam_fpg_frame_step) */

int vslam_aligner_create(int kind, int32_t max_points, int device, vslam_aligner** out);
int vslam_aligner_destroy(vslam_aligner* h);

/* what StereoUVAligner::initialize (:10-69) / UVDAligner::initialize (:11-74) leave in their buffers:
 * moving  n x 3  (_moving)
 * fixed   n x 4  (uL,vL,uR,vR)  StereoUV  |  n x 3 (u,v,depth) UVD            (_fixed)
 * omega   n      scalar of the scalar*I4 information matrix, StereoUV
 *         n x 2  (w_uv, w_depth) of diag(w_uv, w_uv, w_depth), UVD              (_information_matrix_vector)
 * weights_translation n                                                         (_weights_translation)
 * K 3x3 row-major (_camera_calibration_matrix); baseline[3] (_offset_camera_right, ignored for UVD);
 * rows/cols (_number_of_rows/cols_image); minimum_depth (_minimum_reliable_depth_meters). */
int vslam_aligner_upload(vslam_aligner* h, int32_t n, const double* moving, const double* fixed,
                         const double* omega, const double* weights_translation, const double K[9],
                         const double baseline[3], int32_t rows, int32_t cols, double minimum_depth);

/* BaseAligner::linearize(ignore_outliers): stereouv_aligner.cpp:72-187 / uvd_aligner.cpp:77-171.
 * previous_to_current: row-major 3x4 [R|t]. */
int vslam_aligner_linearize(vslam_aligner* h, const double previous_to_current[12], int ignore_outliers,
                            double maximum_error_kernel, vslam_linear_system* system);
/* _errors / _inliers of the last linearize (either may be NULL) */
int vslam_aligner_download(vslam_aligner* h, double* errors, uint8_t* inliers);

/* BaseAligner::oneRound: :190-207 / :174-191 (linearize, damping, 6x6 full-pivot LU, v2t, re-orthonormalise) */
int vslam_aligner_one_round(vslam_aligner* h, const vslam_aligner_parameters* parameters, int ignore_outliers,
                            double previous_to_current[12], vslam_linear_system* system);
/* BaseAligner::converge: :210-264 / :194-248.  information[36] = _information_matrix (may be NULL). */
int vslam_aligner_converge(vslam_aligner* h, const vslam_aligner_parameters* parameters,
                           double previous_to_current[12], vslam_linear_system* system, double* information,
                           int32_t* has_system_converged, int32_t* number_of_rounds);

/* converge() as ONE persistent cooperative kernel (linearize, grid barrier, 6x6 full-pivot LU, v2t update and the
 * convergence logic on the device): no host round trip per Gauss-Newton round.  Same arguments and results as
 * vslam_aligner_converge, bit-identical pose and round count. */
int vslam_aligner_converge_fused(vslam_aligner* h, const vslam_aligner_parameters* parameters,
                                 double previous_to_current[12], vslam_linear_system* system, double* information,
                                 int32_t* has_system_converged, int32_t* number_of_rounds);

/* asynchronous linearize without the host read-back (for callers that time the kernel with CUDA events) */
int vslam_aligner_linearize_async(vslam_aligner* h, const double previous_to_current[12], int ignore_outliers,
                                  double maximum_error_kernel);
int vslam_aligner_read_system(vslam_aligner* h, vslam_linear_system* system);

void* vslam_aligner_stream(vslam_aligner* h);
int vslam_aligner_synchronize(vslam_aligner* h);
int64_t vslam_aligner_launch_count(const vslam_aligner* h);

/* ------------------------------------------------------------------------------------------------------------------
 * what follows the aligner in PoseTracker3D::compute -- the landmark refinement -- and the
 * trajectory wire formats.  Neither sits behind the two plugin classes: PoseTracker3D::_updatePoints
 * (src/position_tracking/pose_tracker_3d.cpp:475-549) calls Landmark::update per point, so a host replaces that LOOP
 * (gather the histories, one call, write the results back), see INTEGRATION.md section 6.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct vslam_landmark_optimizer vslam_landmark_optimizer;

/* Landmark::Measurement (src/types/landmark.h:19-33): frame = index into the pose tables of the call */
typedef struct {
  int32_t frame;
  int32_t reserved;
  double camera_coordinates[3];   /* FramePoint::cameraCoordinatesLeft() */
  double inverse_depth_meters;    /* 1 / camera_coordinates.z */
} vslam_landmark_measurement;

#define VSLAM_LANDMARK_NOT_CONVERGED 0 /* iteration cap reached: state untouched (landmark.cpp:91) */
#define VSLAM_LANDMARK_ADOPTED 1       /* estimate and number of updates replaced (:143-147) */
#define VSLAM_LANDMARK_RESET 2         /* fewer inliers than outliers: average of the measurements (:150-160) */
#define VSLAM_LANDMARK_KEPT 3          /* converged, neither branch taken */

int vslam_landmark_optimizer_create(int32_t max_landmarks, int32_t max_measurements, int32_t max_frames, int device,
                                    vslam_landmark_optimizer** out);
int vslam_landmark_optimizer_destroy(vslam_landmark_optimizer* h);
/* Landmark::update (src/types/landmark.cpp:82-167) for n_landmarks independent landmarks, one warp each.
 * measurement_offsets[n_landmarks + 1]: CSR over `measurements`, the new measurement already appended (:78-79).
 * world_to_camera_left / camera_left_to_world: [n_frames][12] row-major 3x4 (Frame::worldToCameraLeft /
 * cameraLeftToWorld).  world_coordinates[n][3] and number_of_updates[n] are _world_coordinates / _number_of_updates,
 * updated in place exactly where the reference updates them; outcome[n] (VSLAM_LANDMARK_*) and iterations[n] may be
 * NULL.  Parameters: LandmarkParameters::maximum_number_of_iterations / maximum_error_squared_meters
 * (src/types/parameters.h:111-114). */
int vslam_landmark_optimizer_update(vslam_landmark_optimizer* h, int32_t n_landmarks, const int32_t* measurement_offsets,
                                    const vslam_landmark_measurement* measurements, int32_t n_frames,
                                    const double* world_to_camera_left, const double* camera_left_to_world,
                                    uint32_t maximum_number_of_iterations, double maximum_error_squared_meters,
                                    double* world_coordinates, uint32_t* number_of_updates, uint8_t* outcome,
                                    int32_t* iterations);
int64_t vslam_landmark_optimizer_launch_count(const vslam_landmark_optimizer* h);

/* The same refinement over a DEVICE-RESIDENT landmark map: the measurement histories (Landmark::_measurements,
 * src/types/landmark.h:117), world coordinates and update counts stay in HBM, a frame appends ONE measurement per
 * tracked landmark -- what Landmark::update does first (landmark.cpp:71-79) -- and refines it; nothing is re-uploaded.
 * Landmarks are named by the ids vslam_landmark_map_update_frame hands out (0, 1, 2, ... in creation order, like
 * Landmark::_identifier, landmark.cpp:8).  Frames are named by a slot < max_frames chosen by the host (the frame's
 * identifier modulo max_frames while older frames are still referenced by measurements). */
typedef struct vslam_landmark_map vslam_landmark_map;
/* max_measurement_blocks: pool of 32-measurement blocks shared by all landmarks (40 B per measurement);
 * a landmark holds at most 32 * 64 = 2048 measurements */
int vslam_landmark_map_create(int32_t max_landmarks, int32_t max_measurement_blocks, int32_t max_frames, int device,
                              vslam_landmark_map** out);
int vslam_landmark_map_destroy(vslam_landmark_map* h);
/* Frame::setRobotToWorld for an already registered frame (pose graph optimisation moves old frames,
 * src/types/frame.cpp:43-56): later refinements evaluate that frame's measurements with the new poses */
int vslam_landmark_map_set_frame_pose(vslam_landmark_map* h, int32_t frame, const double world_to_camera_left[12],
                                      const double camera_left_to_world[12]);
/* PoseTracker3D::_updatePoints (src/position_tracking/pose_tracker_3d.cpp:475-521) for one frame, as ONE call:
 *  - the frame's poses are stored in slot `frame`;
 *  - n_new landmarks are created (WorldMap::createLandmark -> Landmark::Landmark, landmark.cpp:8-33): landmark i takes the
 *    measurements new_tracks[new_track_offsets[i] .. new_track_offsets[i + 1]) -- its track of framepoints, NEWEST FIRST
 *    as the constructor walks it -- and the average world position new_world[i] the host formed from the framepoints;
 *    new_ids[i] receives its id;
 *  - n_updates existing landmarks (ids unique within the call) receive the measurement
 *    (frame, camera_coordinates[i], 1 / z) and are refined (Landmark::update).  world_out[i][3], updates_out[i],
 *    outcome[i] (VSLAM_LANDMARK_*; may be NULL) and iterations[i] (may be NULL) describe landmark ids[i] after the call.
 * One host -> device copy, two kernels, one device -> host copy. */
int vslam_landmark_map_update_frame(vslam_landmark_map* h, int32_t frame, const double world_to_camera_left[12],
                                    const double camera_left_to_world[12], int32_t n_updates, const int32_t* ids,
                                    const double* camera_coordinates, int32_t n_new, const int32_t* new_track_offsets,
                                    const vslam_landmark_measurement* new_tracks, const double* new_world,
                                    uint32_t maximum_number_of_iterations, double maximum_error_squared_meters,
                                    double* world_out, uint32_t* updates_out, uint8_t* outcome, int32_t* iterations,
                                    int32_t* new_ids);
/* state of any landmarks (e.g. before writing the map): world[n][3], number_of_updates[n], n_measurements[n]; NULLs skipped */
int vslam_landmark_map_get(vslam_landmark_map* h, int32_t n, const int32_t* ids, double* world, uint32_t* number_of_updates,
                           int32_t* n_measurements);
int32_t vslam_landmark_map_size(const vslam_landmark_map* h);
int64_t vslam_landmark_map_launch_count(const vslam_landmark_map* h);

/* WorldMap::writeTrajectoryKITTI (src/types/world_map.cpp:183-216) / writeTrajectoryTUM (:218-252): one text line per
 * frame, std::fixed with 9 decimals, every value followed by a blank.  KITTI: the 12 values of robot_to_world row by
 * row.  TUM: timestamp, translation, orientation quaternion x y z w.  The format_ calls return the line length
 * (snprintf semantics); write_trajectory overwrites `filename`. */
#define VSLAM_TRAJECTORY_KITTI 0
#define VSLAM_TRAJECTORY_TUM 1
int32_t vslam_format_trajectory_kitti(const double robot_to_world[12], char* line, int32_t capacity);
int32_t vslam_format_trajectory_tum(double timestamp_seconds, const double robot_to_world[12], char* line, int32_t capacity);
int vslam_write_trajectory(const char* filename, int format, int32_t n_frames, const double* robot_to_world,
                           const double* timestamps_seconds);

/* the 3x3 complete-pivoting LU solve of Landmark::update (exposed for tests) */
void vslam_solve3(const double A[9], const double rhs[3], double x[3]);

/* host-side 6x6 helpers used by one_round (exposed for tests): complete-pivoting LU solve; srrg_core::v2t */
void vslam_solve6(const double A[36], const double rhs[6], double x[6]);
void vslam_v2t(const double v[6], double T[12]);

#ifdef __cplusplus
}
#endif
#endif /* VSLAM_B200_H */
