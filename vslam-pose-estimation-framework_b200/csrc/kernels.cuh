// kernels.cuh -- launch interface of the framepoint-generation and aligner kernels (sm_100a).
#pragma once

#include <cuda.h>   // CUtensorMap (type only; the encoder is resolved through cudaGetDriverEntryPoint)

#include "common.cuh"

namespace vslam {

struct RegionTable {
  Region r[kMaxRegions];
  int threshold[kMaxRegions];   // rint(detector threshold), clamped to [0, 255] like cv::FAST
};

struct StereoParams {
  double fx, fy, cx, cy, bx;
  double max_matching_distance;   // maximum_matching_distance_triangulation
  double min_disparity;           // minimum_disparity_pixels
  int target_keypoints;           // _target_number_of_keypoints
  int localizing;                 // frame->status() == Frame::Localizing
};

struct TrackedPoint {   // == vslam_tracked_point
  int32_t row, col, has_previous, reserved;
  double disparity, distance;
};

struct FramePointRecord {   // == vslam_framepoint
  int32_t index_left, index_right;
  float xl, yl, xr, yr;
  int32_t distance, epipolar_offset;
  double camera[3];
};

struct PreviousPoint {   // == vslam_previous_point (128 bytes)
  double camera[3];
  double world[3];
  uint8_t descriptor_left[32], descriptor_right[32];
  int32_t epipolar_offset, has_landmark;
  float keypoint_size;
  int32_t reserved;
};

struct TrackRecord {   // == vslam_track (88 bytes)
  int32_t index_previous, index_left, index_right;
  float xl, yl, xr, yr;
  int32_t distance, epipolar_offset;
  float projection_left[2], projection_right[2], projection_right_corrected[2];
  int32_t reserved;
  double camera[3];
};

struct RecoveredRecord {   // == vslam_recovered_point (112 bytes)
  int32_t index_lost, distance;
  float xl, yl, xr, yr;
  double camera[3];
  uint8_t descriptor_left[32], descriptor_right[32];
};

struct TrackParams {
  double T[12];                    // camera_left_previous_in_current, row-major 3x4
  int by_appearance;               // track_by_appearance_
  int distance_pixels;             // _projection_tracking_distance_pixels
  double max_distance_tracking;    // _maximum_descriptor_distance_tracking
};

// Device-resident inputs and counters of the fused tracked frame (vslam_fpg_frame_step).  A captured CUDA graph
// re-launches every kernel of the frame with the same arguments, so everything that changes from frame to frame --
// the motion prior and the point counts the stages hand to each other -- lives here instead of in kernel arguments.
struct FrameStepState {
  double T_prior[12];      // camera_left_previous_in_current of this frame (copy node of the graph, from pinned memory)
  int32_t frame_id;        // (same copy) the host's number of this frame, echoed into the result header when it is complete
  int32_t ticket;          // (same copy: 0) blocks of the last kernel that have published their share
  int32_t thresholds[kMaxRegions];   // (same copy) the FAST thresholds of this frame, per detector region
  int32_t n_previous;      // points() of the previous frame held in `previous` (written by frame_assemble_kernel)
  int32_t n_kept;          // tracks that survive _prunePoints (frame_prune_kernel)
  int32_t inliers_only;    // the branch of pose_tracker_3d.cpp:441 taken by frame_prune_kernel
  int32_t overflow;        // != 0: more points than the fused path holds (the stepwise calls have no such limit)
};

struct TrackScratch {
  int4* tentative;       // [n_previous] {sorted left feature | -1, sorted right feature | -1, distance, status}
  int4* final;           // [n_previous] per track, in order: {previous point, left feature, right feature, distance}
  int32_t* claim_l;      // [cap] lowest previous-point index that consumes the left feature
  int32_t* claim_r;      // [cap]
  int32_t* stats;        // [4] {tracks, lost, tracked landmarks, accumulated distance}
};

struct RecoverParams {
  double W[12];                    // world_to_camera_left
  double min_depth, max_depth;     // minimum_depth_meters / maximum_depth_meters
  double max_distance_tracking;
  const int8_t* brief_tests;       // non-null: BRIEF-32 on the box-sum images instead of rBRIEF on the blurred ones
};

// all launches are asynchronous on `stream`; image ranges are [first_image, first_image + n_images)
// dense host-layout images (row stride `stride` bytes, any alignment) -> pitched device images [pair][side][row][pitch]
// clear (optional): n_clear i32 zeroed by the kernel (the raw FAST counters of a single frame)
void launch_repitch(const Geometry& g, const uint8_t* left, const uint8_t* right, int stride, uint8_t* image,
                    int n_pairs, cudaStream_t stream, int32_t* clear = nullptr, int n_clear = 0);
// `image_map`: TMA descriptor of b.image (all images of the handle) with the FAST tile as box
bool make_fast_tensor_map(const Geometry& g, const uint8_t* images, int n_images, CUtensorMap* out);
bool make_image_tensor_map(const Geometry& g, const uint8_t* images, int n_images, int box_w, int box_h, CUtensorMap* out);
// device_thresholds (optional): [n_regions] i32 in device memory, read instead of rt.threshold
void launch_fast(const Geometry& g, const RegionTable& rt, const Buffers& b, const CUtensorMap& image_map, int first_image,
                 int n_images, cudaStream_t stream, const int32_t* device_thresholds = nullptr,
                 bool counts_cleared = false, bool mask_cleared = false);
void launch_compact(const Geometry& g, const Buffers& b, int first_image, int n_images, cudaStream_t stream);
// layout of the packed features of pair 0 (pack_features_kernel): per side [n] u32 | [n] u8 padded to 16 | [n][32]
__host__ __device__ inline size_t feature_pack_desc_offset(int n) { return (5 * (size_t)n + 15) / 16 * 16; }
__host__ __device__ inline size_t feature_pack_bytes(int n) { return feature_pack_desc_offset(n) + 32 * (size_t)n; }
// positions, FAST responses and descriptors of both sides of pair 0, packed for one device-to-host copy
void launch_pack_features(const Geometry& g, const Buffers& b, uint8_t* out, cudaStream_t stream);
// FAST response of the kept keypoints of one image -> kp_score (on demand)
void launch_score(const Geometry& g, const Buffers& b, int image, cudaStream_t stream);
// `image_map`: TMA descriptor of b.image (all images of the handle) with the blur tile as box
bool make_blur_tensor_map(const Geometry& g, const uint8_t* images, int n_images, CUtensorMap* out);
void launch_blur(const Geometry& g, const Buffers& b, const CUtensorMap& image_map, int first_image, int n_images,
                 cudaStream_t stream);
// TMA descriptor of a blurred scratch buffer [n_images][rows][pitch]; false when the driver cannot encode it
bool make_blurred_tensor_map(const Geometry& g, const uint8_t* blurred, int n_images, CUtensorMap* out);
// image `scratch_first` of `blurred_map` (the lane's blurred scratch) is image `first_image` of the batch
void launch_describe(const Geometry& g, const Buffers& b, const CUtensorMap& blurred_map, int first_image, int n_images,
                     cudaStream_t stream, int scratch_first = 0);
// one launch per epipolar offset (pass index -> offset 0,+1,-1,+2,...)
void launch_match(const Geometry& g, const StereoParams& sp, const Buffers& b, int first_pair, int n_pairs, int pass,
                  int epipolar_offset, cudaStream_t stream);
void launch_select(const Geometry& g, const StereoParams& sp, const Buffers& b, int first_pair, int n_pairs,
                   int n_passes, const TrackedPoint* tracked, int n_tracked, FramePointRecord* out,
                   int out_capacity_per_pair, bool generic, cudaStream_t stream,
                   const int32_t* n_tracked_device = nullptr, int mode = 0, int32_t* bins = nullptr);
// fused frame: the strip kernel in two halves (kSelectReplay beside the aligner -> bins[rows_bin * cols_bin + 1],
// kSelectMerge after the prune); only where the strip kernel serves the bin grid
enum { kSelectAll = 0, kSelectReplay = 1, kSelectMerge = 2 };
bool select_strips_available(const Geometry& g);
void launch_emit_matches(const Geometry& g, const StereoParams& sp, const Buffers& b, int pair, int n_passes,
                         FramePointRecord* out, int out_capacity, int32_t* n_out, cudaStream_t stream);
int kernels_per_match_pass();
// rBRIEF-256 at arbitrary interior pixels of `n_images` blurred images: image i reads xy[i * stride .. + n[i]) and
// writes desc[(i * stride + k) * 32]
void launch_describe_at(const Geometry& g, const uint8_t* blurred, const uint32_t* xy, const int32_t* n, uint8_t* desc,
                        int stride, int n_images, cudaStream_t stream);
// BRIEF-32 (cv::xfeatures2d::BriefDescriptorExtractor, base_framepoint_generator.cpp:186): 9x9 box sums of the images
// (u16 [n_images][rows][pitch]) and the 256 box-sum comparisons of a supplied test table (device, 256 x 4 int8:
// y0, x0, y1, x1) at the keypoints xy[i * stride .. + n[i]) of image i
void launch_box9(const Geometry& g, const uint8_t* image, uint16_t* boxsum, int n_images, cudaStream_t stream);
void launch_describe_brief(const Geometry& g, const uint16_t* boxsum, const int8_t* tests, const uint32_t* xy,
                           const int32_t* n, uint8_t* desc, int stride, int n_images, cudaStream_t stream);
// StereoFramePointGenerator::track for pair `pair`: two launches (parallel search, ordered resolution).  Marks the
// consumed features in pruned_l / consumed_r, writes tracks / lost (ordered), the bin pre-load records and stats.
void launch_track(const Geometry& g, const StereoParams& sp, const Buffers& b, int pair, const PreviousPoint* previous,
                  int n_previous, const TrackParams& tp, const TrackScratch& scratch, TrackRecord* tracks,
                  int32_t* lost, TrackedPoint* tracked, cudaStream_t stream, const FrameStepState* step = nullptr);
// StereoFramePointGenerator::recoverPoints for pair `pair` (blurred images at `blurred`): three launches
void launch_recover(const Geometry& g, const StereoParams& sp, const Buffers& b, int pair, const uint8_t* blurred,
                    const PreviousPoint* lost, int n_lost, const RecoverParams& rp, uint32_t* xy, int32_t* n_xy,
                    uint8_t* desc, RecoveredRecord* out, int32_t* n_out, cudaStream_t stream);

// ---- aligner ----
struct AlignerBuffers {
  const double* moving;   // SoA: [3][stride]
  const double* fixed;    // SoA: [4 or 3][stride]
  const double* omega;    // SoA: [1 or 2][stride]
  const double* wt;       // [stride]
  double* errors;         // [stride]
  uint8_t* inliers;       // [stride]
  double* partials;       // [grid][32]
  double* system;         // [32] : H upper triangle (21), b (6), total error, inliers
  unsigned int* ticket;   // [1]
  int stride;
};

struct AlignerCamera {
  double K[9];
  double baseline[3];
  double rows, cols;
  double min_depth;
};

// parameters and device-resident state of the fused Gauss-Newton kernel
struct GnParams {
  double error_delta;   // error_delta_for_convergence
  double kernel;        // maximum_error_kernel
  double damping;
  int max_iterations;   // maximum_number_of_iterations
  int inlier_gate;      // minimum_number_of_inliers (StereoUV) / 100 (UVD)
};

struct GnControl {
  double T[12];                  // previous_to_current, in/out
  double H[36];                  // damped H of the last round (-> _information_matrix)
  double total_error_previous;
  int rounds, phase, iteration, ignore, converged, done;
};

// fused frame: StereoUVAligner::initialize (reference src/aligners/stereouv_aligner.cpp:26-64, the branch without a
// landmark estimate) as the head of the cluster Gauss-Newton kernel -- thread k reads track k and its previous point
// instead of planes another kernel would have to write; the control block of converge() starts from the motion prior.
// the landmark estimate of a point of the previous frame (stereouv_aligner.cpp:43-51): when information_scale != 0 the
// aligner moves `camera` (previous->cameraCoordinatesLeftLandmark()) instead of the point's own camera coordinates and
// scales its information by information_scale (the host's 1 + log(landmark->numberOfUpdates())).  == vslam_landmark_estimate
struct LandmarkEstimate {
  double camera[3];
  double information_scale;
};

struct FrameFill {
  const TrackRecord* tracks;       // [cap] track() output; nullptr: the correspondences come from AlignerBuffers
  const PreviousPoint* previous;   // points() of the previous frame
  const LandmarkEstimate* estimates;   // [cap] per point of the previous frame (vslam_fpg_frame_step_set_landmark_estimates)
  int32_t* track_length;           // [cap] per track: trackLength() of its previous point (for the frame's points())
  FrameStepState* state;           // T_prior in; overflow, n_kept, inliers_only out
  double max_reliable_depth;       // _maximum_reliable_depth_meters (slam_assembly.cpp:70)
  int inverse_depth_weight;        // enable_inverse_depth_as_information
};

// fused frame: PoseTracker3D::_prunePoints (reference src/position_tracking/pose_tracker_3d.cpp:437-472) as the tail of the
// cluster Gauss-Newton kernel -- the errors / inliers of the last round are still in the threads' registers.  Record k of
// the bin pre-load belongs to correspondence k; the kept records are compacted in place, in order.
struct FramePrune {
  TrackedPoint* tracked;   // [cap] bin pre-load records of track(); nullptr: no prune
  int32_t* kept_pos;       // [cap] per track: position among the surviving tracks or -1
  FrameStepState* state;   // n_kept, inliers_only
  double error_kernel;     // maximum_error_kernel
};

int aligner_grid(int n, int resident_blocks);
int converge_max_blocks_per_sm(int kind);
cudaError_t launch_converge(int kind, int n, const AlignerBuffers& b, const AlignerCamera& cam, const GnParams& p,
                            GnControl* ctl, int grid, cudaStream_t stream);
void launch_linearize(int kind, int n, const AlignerBuffers& b, const AlignerCamera& cam, const double T[12],
                      int ignore_outliers, double kernel, int grid, cudaStream_t stream);
// StereoUV converge of the fused tracked frame: ONE thread-block cluster of `cluster_blocks` CTAs (the same blocks as
// launch_converge's, so the pose sequence is bit-identical), the correspondence count read from *n_device (0: the
// kernel leaves the control block as it is, rounds = 0).  frame_step_cluster_blocks(): the largest cluster this device
// schedules for the kernel (16 with the non-portable opt-in, else 8, 0 when none); capacity = blocks x 256.
int frame_step_cluster_blocks();
cudaError_t launch_converge_frame(const AlignerBuffers& b, const AlignerCamera& cam, const GnParams& p, GnControl* ctl,
                                  const int32_t* n_device, int cluster_blocks, const FrameFill& fill, const FramePrune& prune,
                                  cudaStream_t stream);
// one CTA per stereo pair: StereoUV initialize + linearize of the pair's new framepoints against themselves
void launch_linearize_pairs(const FramePointRecord* records, int record_stride, const int32_t* n_out, int n_pairs,
                            const AlignerCamera& cam, const double T[12], int ignore_outliers, double kernel,
                            double max_reliable_depth, int inverse_depth_weight, double* systems, double* errors,
                            uint8_t* inliers, cudaStream_t stream);

// Landmark::Measurement (reference src/types/landmark.h:19-33) as the device reads it; == vslam_landmark_measurement
struct LandmarkMeasurement {
  int32_t frame;                   // index into the pose tables
  int32_t reserved;
  double camera_coordinates[3];
  double inverse_depth_meters;
};
// PoseTracker3D::_prunePoints on the bin pre-load records of the last track() (at most 8192), ordered, in place
void launch_prune_tracked(TrackedPoint* tracked, int n, const double* errors, const uint8_t* inliers, int inliers_only,
                          double error_cap, cudaStream_t stream);
// ---- fused tracked frame (frame_step.cu) ----
// what the frame publishes into the handle's pinned result block, written by the device through the mapped pointer
struct FrameStepHeader {
  int32_t stats[4];        // track(): {tracks, lost, tracked landmarks, accumulated descriptor distance}
  int32_t n_kept, inliers_only, overflow, error_flag;
  int32_t n_out[2];        // compute(): {new framepoints, matches}
  int32_t n_previous;      // points the frame was tracked against
  int32_t n_points;        // points() of this frame = n_kept + new framepoints
  GnControl ctl;           // pose, damped H, rounds, converged
  double system[32];       // last linearisation: H upper triangle, b, total error, inliers
  int32_t done_frame;      // == the frame_id of the call once EVERYTHING of the frame is in this block (written last)
  int32_t reserved;
};
struct FrameStepBuffers {
  FrameStepState* state;
  PreviousPoint* previous;          // [cap] points() of the previous frame, replaced in place at the end of the frame
  const TrackRecord* tracks;        // [cap] track() output
  const int32_t* lost;              // [cap]
  TrackedPoint* tracked;            // [cap] bin pre-load records (pruned in place)
  const int32_t* stats;             // track() stats
  int32_t* track_length;            // [cap] per track: trackLength() of its previous point
  int32_t* kept_pos;                // [cap] per track: position among the surviving tracks or -1
  const FramePointRecord* points;   // [out_cap] compute() output
  const int32_t* n_out;             // {new framepoints, matches}
  const int32_t* error_flag;
  const uint8_t* desc;              // descriptors of the frame [2][g.cap][32]
  GnControl* ctl;
  AlignerBuffers aligner;           // StereoUV planes with stride = cap
  int cap, out_cap;
  // pinned, device-mapped result block
  FrameStepHeader* h_header;
  TrackRecord* h_tracks;            // [cap] surviving tracks, in order
  uint8_t* h_kept;                  // [cap] per track of track()
  double* h_errors;                 // [cap]
  uint8_t* h_inliers;               // [cap]
  int32_t* h_lost;                  // [cap]
  FramePointRecord* h_points;       // [out_cap]
  PreviousPoint* h_frame_points;    // [cap] points() of this frame (the next frame's previous points)
  LandmarkEstimate* estimates;      // [cap] landmark estimates of points(): a new frame's points have none until the host says so
  // detection status of the frame ({error flag, pad, n_desc[2]} and the raw FAST counts per region), mirrored into the
  // handle's pinned status words by the last block of frame_assemble_kernel (no copy node at the end of the graph)
  const int32_t* d_status;          // [4]
  const int32_t* d_raw_count;       // [2 * n_regions]
  int32_t* h_status;                // pinned, device-visible
  int32_t* h_counts;                // pinned, device-visible
  int n_regions;
};
struct FrameStepParams {
  double max_reliable_depth;        // _maximum_reliable_depth_meters of the aligner (slam_assembly.cpp:70)
  int inverse_depth_weight;         // enable_inverse_depth_as_information
  double error_kernel;              // maximum_error_kernel (the prune rule)
  int min_track_length;             // minimum_track_length_for_landmark_creation: has_landmark of the assembled points
  int publish_frame_points;         // copy points() of the frame (128 B each) to the host as well
};
// StereoUVAligner::initialize over the tracks (stereouv_aligner.cpp:26-64, the branch without a landmark estimate) and the
// control block of converge(); then, after the cluster kernel: _prunePoints; after select: points() of the frame
enum { kAssembleAll = 0, kAssembleTracks = 1, kAssembleRest = 2 };   // blocks of one frame_assemble launch
void launch_frame_assemble(const Geometry& g, const FrameStepBuffers& f, const FrameStepParams& p, int part,
                           cudaStream_t stream);

// device-resident landmark map (vslam_landmark_map): per landmark a chain of 32-measurement blocks, its world
// coordinates and update count; per frame slot the two poses every measurement of that frame is evaluated with
struct LandmarkMapBuffers {
  int32_t* count;                  // [max_landmarks] measurements held
  int32_t* table;                  // [max_landmarks][blocks_per_landmark] block ids
  LandmarkMeasurement* blocks;     // [max_blocks][32]
  int32_t* next_block;             // pool cursor
  int32_t* error;                  // set when the pool or a landmark's table is exhausted
  double* world;                   // [max_landmarks][3]
  uint32_t* updates;               // [max_landmarks]
  const double* world_to_camera;   // [max_frames][12]
  const double* camera_to_world;   // [max_frames][12]
  int blocks_per_landmark, max_blocks;
};
void launch_landmark_create(const LandmarkMapBuffers& b, int n_new, int first_id, const int32_t* track_offsets,
                            const LandmarkMeasurement* tracks, const double* world_init, cudaStream_t stream);
void launch_landmark_map_update(const LandmarkMapBuffers& b, int n, const int32_t* ids, const double* camera_coordinates,
                                int frame, uint32_t max_iterations, double max_err2, uint8_t* outcome, int32_t* iterations,
                                double* world_out, uint32_t* updates_out, cudaStream_t stream);
// one warp per landmark: Landmark::update's Gauss-Newton (landmark.cpp:82-167); poses are row-major 3x4 per frame
void launch_landmark_update(int n_landmarks, const int32_t* offsets, const LandmarkMeasurement* measurements,
                            const double* world_to_camera, const double* camera_to_world, uint32_t max_iterations,
                            double max_err2, double* world, uint32_t* number_of_updates, uint8_t* outcome,
                            int32_t* iterations, cudaStream_t stream);

}  // namespace vslam
