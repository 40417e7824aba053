// landmark_api.cu -- C ABI of SURVEY.md 8f row 4 (include/vslam_b200.h): the batched landmark refinement
// (Landmark::update, reference src/types/landmark.cpp:66-152, driven by PoseTracker3D::_updatePoints,
// src/position_tracking/pose_tracker_3d.cpp:475-549) and the trajectory wire formats
// (WorldMap::writeTrajectoryKITTI / writeTrajectoryTUM, src/types/world_map.cpp:183-252).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>

#include "../../include/vslam_b200.h"
#include "api_common.h"
#include "gn_math.h"
#include "host_math.h"
#include "kernels.cuh"

using namespace vslam;

static_assert(sizeof(vslam_landmark_measurement) == sizeof(LandmarkMeasurement) && sizeof(LandmarkMeasurement) == 40,
              "measurement layout");

struct vslam_landmark_optimizer {
  int device = 0;
  int32_t max_landmarks = 0, max_measurements = 0, max_frames = 0;
  cudaStream_t stream = nullptr;
  int32_t* d_offsets = nullptr;
  LandmarkMeasurement* d_measurements = nullptr;
  double* d_poses = nullptr;            // [2][max_frames][12]: world_to_camera_left, then camera_left_to_world
  double* d_world = nullptr;
  uint32_t* d_updates = nullptr;
  uint8_t* d_outcome = nullptr;
  int32_t* d_iterations = nullptr;
  int64_t launches = 0;
};

// Device-resident landmark map.  One pinned staging block per direction:
//   IN  : [poses 24 f64][ids n i32 (padded)][camera coordinates 3 n f64][new offsets (n_new + 1) i32 (padded)]
//         [new tracks: measurements][new world 3 n_new f64]
//   OUT : [error i32, pad][world 3 n f64][updates n u32][iterations n i32][outcome n u8]
struct vslam_landmark_map {
  int device = 0;
  int32_t max_landmarks = 0, max_blocks = 0, max_frames = 0;
  int32_t n_landmarks = 0;
  cudaStream_t stream = nullptr;
  LandmarkMapBuffers b = {};
  double* d_poses = nullptr;      // [2][max_frames][12]
  uint8_t* d_in = nullptr;
  uint8_t* h_in = nullptr;        // pinned
  uint8_t* d_out = nullptr;
  uint8_t* h_out = nullptr;       // pinned
  size_t in_capacity = 0, out_capacity = 0;
  int64_t launches = 0;
};

namespace {
constexpr int kBlocksPerLandmark = 64;
size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

int ensure_io(vslam_landmark_map* h, size_t in_bytes, size_t out_bytes) {
  if (in_bytes > h->in_capacity) {
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_in);
    cudaFreeHost(h->h_in);
    h->d_in = h->h_in = nullptr;
    h->in_capacity = 0;
    const size_t cap = std::max<size_t>(2 * in_bytes, 1 << 16);
    CUDA_TRY(cudaMalloc((void**)&h->d_in, cap));
    CUDA_TRY(cudaMallocHost((void**)&h->h_in, cap));
    h->in_capacity = cap;
  }
  if (out_bytes > h->out_capacity) {
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_out);
    cudaFreeHost(h->h_out);
    h->d_out = h->h_out = nullptr;
    h->out_capacity = 0;
    const size_t cap = std::max<size_t>(2 * out_bytes, 1 << 16);
    CUDA_TRY(cudaMalloc((void**)&h->d_out, cap));
    CUDA_TRY(cudaMallocHost((void**)&h->h_out, cap));
    h->out_capacity = cap;
  }
  return VSLAM_OK;
}
}  // namespace

extern "C" {

int vslam_landmark_map_create(int32_t max_landmarks, int32_t max_measurement_blocks, int32_t max_frames, int device,
                              vslam_landmark_map** out) {
  if (!out) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  if (max_landmarks < 1 || max_measurement_blocks < 1 || max_frames < 1)
    return fail(VSLAM_ERR_INVALID_ARGUMENT, "capacities must be positive");
  int rc = require_device(device);
  if (rc) return rc;
  vslam_landmark_map* h = new vslam_landmark_map();
  h->device = device;
  h->max_landmarks = max_landmarks;
  h->max_blocks = max_measurement_blocks;
  h->max_frames = max_frames;
  bool ok = cudaSetDevice(device) == cudaSuccess && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
  auto dalloc = [&](void** p, size_t bytes) { ok = ok && cudaMalloc(p, bytes) == cudaSuccess; };
  LandmarkMapBuffers& b = h->b;
  dalloc((void**)&b.count, sizeof(int32_t) * (size_t)max_landmarks);
  dalloc((void**)&b.table, sizeof(int32_t) * (size_t)max_landmarks * kBlocksPerLandmark);
  dalloc((void**)&b.blocks, sizeof(LandmarkMeasurement) * 32 * (size_t)max_measurement_blocks);
  dalloc((void**)&b.next_block, 2 * sizeof(int32_t));
  dalloc((void**)&b.world, sizeof(double) * 3 * (size_t)max_landmarks);
  dalloc((void**)&b.updates, sizeof(uint32_t) * (size_t)max_landmarks);
  dalloc((void**)&h->d_poses, sizeof(double) * 24 * (size_t)max_frames);
  if (ok) {
    b.error = b.next_block + 1;
    b.world_to_camera = h->d_poses;
    b.camera_to_world = h->d_poses + 12 * (size_t)max_frames;
    b.blocks_per_landmark = kBlocksPerLandmark;
    b.max_blocks = max_measurement_blocks;
    ok = cudaMemset(b.next_block, 0, 2 * sizeof(int32_t)) == cudaSuccess &&
         cudaMemset(b.count, 0, sizeof(int32_t) * (size_t)max_landmarks) == cudaSuccess;
  }
  if (!ok) {
    const cudaError_t e = cudaGetLastError();
    vslam_landmark_map_destroy(h);
    return fail(VSLAM_ERR_CUDA, "landmark map allocation failed: %s", cudaGetErrorString(e));
  }
  *out = h;
  return VSLAM_OK;
}

int vslam_landmark_map_destroy(vslam_landmark_map* h) {
  if (!h) return VSLAM_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  cudaFree(h->b.count); cudaFree(h->b.table); cudaFree(h->b.blocks); cudaFree(h->b.next_block);
  cudaFree(h->b.world); cudaFree(h->b.updates); cudaFree(h->d_poses);
  cudaFree(h->d_in); cudaFree(h->d_out);
  cudaFreeHost(h->h_in); cudaFreeHost(h->h_out);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return VSLAM_OK;
}

int vslam_landmark_map_set_frame_pose(vslam_landmark_map* h, int32_t frame, const double w2c[12], const double c2w[12]) {
  if (!h || !w2c || !c2w) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  if (frame < 0 || frame >= h->max_frames) return fail(VSLAM_ERR_CAPACITY, "frame slot %d outside [0, %d)", frame, h->max_frames);
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemcpyAsync(h->d_poses + 12 * (size_t)frame, w2c, sizeof(double) * 12, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->d_poses + 12 * ((size_t)h->max_frames + frame), c2w, sizeof(double) * 12, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return VSLAM_OK;
}

int vslam_landmark_map_update_frame(vslam_landmark_map* h, int32_t frame, const double w2c[12], const double c2w[12],
                                    int32_t n_updates, const int32_t* ids, const double* camera_coordinates, int32_t n_new,
                                    const int32_t* new_track_offsets, const vslam_landmark_measurement* new_tracks,
                                    const double* new_world, uint32_t maximum_number_of_iterations,
                                    double maximum_error_squared_meters, double* world_out, uint32_t* updates_out,
                                    uint8_t* outcome, int32_t* iterations, int32_t* new_ids) {
  VSLAM_NVTX("vslam_landmark_map_update_frame [PoseTracker3D::_updatePoints]");
  if (!h || !w2c || !c2w) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  if (frame < 0 || frame >= h->max_frames) return fail(VSLAM_ERR_CAPACITY, "frame slot %d outside [0, %d)", frame, h->max_frames);
  if (n_updates < 0 || n_new < 0) return fail(VSLAM_ERR_INVALID_ARGUMENT, "negative count");
  if (n_updates && (!ids || !camera_coordinates || !world_out || !updates_out)) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null update argument");
  if (n_new && (!new_track_offsets || !new_tracks || !new_world || !new_ids)) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null creation argument");
  if (h->n_landmarks + (int64_t)n_new > h->max_landmarks)
    return fail(VSLAM_ERR_CAPACITY, "%d + %d landmarks exceed the capacity %d", h->n_landmarks, n_new, h->max_landmarks);
  for (int32_t i = 0; i < n_updates; ++i)
    if (ids[i] < 0 || ids[i] >= h->n_landmarks) return fail(VSLAM_ERR_INVALID_ARGUMENT, "unknown landmark id %d", ids[i]);
  int32_t n_track = 0;
  if (n_new) {
    if (new_track_offsets[0] != 0) return fail(VSLAM_ERR_INVALID_ARGUMENT, "new_track_offsets[0] must be 0");
    for (int32_t i = 0; i < n_new; ++i)
      if (new_track_offsets[i + 1] <= new_track_offsets[i]) return fail(VSLAM_ERR_INVALID_ARGUMENT, "new landmark %d has no measurement", i);
    n_track = new_track_offsets[n_new];
    for (int32_t m = 0; m < n_track; ++m)
      if (new_tracks[m].frame < 0 || new_tracks[m].frame >= h->max_frames)
        return fail(VSLAM_ERR_INVALID_ARGUMENT, "track measurement %d names frame slot %d", m, new_tracks[m].frame);
  }
  CUDA_TRY(cudaSetDevice(h->device));
  // ---- pack the inputs into one pinned block
  const size_t o_ids = 24 * sizeof(double);
  const size_t o_cam = o_ids + align16(sizeof(int32_t) * (size_t)n_updates);
  const size_t o_off = o_cam + sizeof(double) * 3 * (size_t)n_updates;
  const size_t o_trk = o_off + align16(sizeof(int32_t) * ((size_t)n_new + 1));
  const size_t o_wld = o_trk + sizeof(LandmarkMeasurement) * (size_t)n_track;
  const size_t in_bytes = o_wld + sizeof(double) * 3 * (size_t)n_new;
  const size_t p_world = 16;
  const size_t p_upd = p_world + sizeof(double) * 3 * (size_t)n_updates;
  const size_t p_it = p_upd + align16(sizeof(uint32_t) * (size_t)n_updates);
  const size_t p_out = p_it + align16(sizeof(int32_t) * (size_t)n_updates);
  const size_t out_bytes = p_out + align16((size_t)n_updates);
  int rc = ensure_io(h, in_bytes, out_bytes);
  if (rc) return rc;
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  std::memcpy(h->h_in, w2c, 12 * sizeof(double));
  std::memcpy(h->h_in + 12 * sizeof(double), c2w, 12 * sizeof(double));
  if (n_updates) {
    std::memcpy(h->h_in + o_ids, ids, sizeof(int32_t) * (size_t)n_updates);
    std::memcpy(h->h_in + o_cam, camera_coordinates, sizeof(double) * 3 * (size_t)n_updates);
  }
  if (n_new) {
    std::memcpy(h->h_in + o_off, new_track_offsets, sizeof(int32_t) * ((size_t)n_new + 1));
    std::memcpy(h->h_in + o_trk, new_tracks, sizeof(LandmarkMeasurement) * (size_t)n_track);
    std::memcpy(h->h_in + o_wld, new_world, sizeof(double) * 3 * (size_t)n_new);
  }
  cudaStream_t s = h->stream;
  CUDA_TRY(cudaMemcpyAsync(h->d_in, h->h_in, in_bytes, cudaMemcpyHostToDevice, s));
  // the frame's poses: device -> device out of the block just uploaded
  CUDA_TRY(cudaMemcpyAsync(h->d_poses + 12 * (size_t)frame, h->d_in, sizeof(double) * 12, cudaMemcpyDeviceToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(h->d_poses + 12 * ((size_t)h->max_frames + frame), h->d_in + 12 * sizeof(double), sizeof(double) * 12,
                           cudaMemcpyDeviceToDevice, s));
  // updates first: a landmark is never created and updated by the same frame (pose_tracker_3d.cpp:499-511)
  launch_landmark_map_update(h->b, n_updates, reinterpret_cast<const int32_t*>(h->d_in + o_ids),
                             reinterpret_cast<const double*>(h->d_in + o_cam), frame, maximum_number_of_iterations,
                             maximum_error_squared_meters, h->d_out + p_out, reinterpret_cast<int32_t*>(h->d_out + p_it),
                             reinterpret_cast<double*>(h->d_out + p_world), reinterpret_cast<uint32_t*>(h->d_out + p_upd), s);
  launch_landmark_create(h->b, n_new, h->n_landmarks, reinterpret_cast<const int32_t*>(h->d_in + o_off),
                         reinterpret_cast<const LandmarkMeasurement*>(h->d_in + o_trk),
                         reinterpret_cast<const double*>(h->d_in + o_wld), s);
  h->launches += (n_updates > 0) + (n_new > 0);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(h->d_out, h->b.error, sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(h->h_out, h->d_out, out_bytes, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (*reinterpret_cast<const int32_t*>(h->h_out))
    return fail(VSLAM_ERR_CAPACITY, "the measurement block pool (%d blocks) or a landmark's history (%d measurements) is full",
                h->max_blocks, 32 * kBlocksPerLandmark);
  if (n_updates) {
    std::memcpy(world_out, h->h_out + p_world, sizeof(double) * 3 * (size_t)n_updates);
    std::memcpy(updates_out, h->h_out + p_upd, sizeof(uint32_t) * (size_t)n_updates);
    if (iterations) std::memcpy(iterations, h->h_out + p_it, sizeof(int32_t) * (size_t)n_updates);
    if (outcome) std::memcpy(outcome, h->h_out + p_out, (size_t)n_updates);
  }
  for (int32_t i = 0; i < n_new; ++i) new_ids[i] = h->n_landmarks + i;
  h->n_landmarks += n_new;
  return VSLAM_OK;
}

int vslam_landmark_map_get(vslam_landmark_map* h, int32_t n, const int32_t* ids, double* world, uint32_t* number_of_updates,
                           int32_t* n_measurements) {
  if (!h || n < 0 || (n && !ids)) return fail(VSLAM_ERR_INVALID_ARGUMENT, "bad arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  for (int32_t i = 0; i < n; ++i) {     // a bookkeeping call (map export), not a per-frame one: plain copies
    if (ids[i] < 0 || ids[i] >= h->n_landmarks) return fail(VSLAM_ERR_INVALID_ARGUMENT, "unknown landmark id %d", ids[i]);
    if (world) CUDA_TRY(cudaMemcpy(world + 3 * (size_t)i, h->b.world + 3 * (size_t)ids[i], 3 * sizeof(double), cudaMemcpyDeviceToHost));
    if (number_of_updates) CUDA_TRY(cudaMemcpy(number_of_updates + i, h->b.updates + ids[i], sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (n_measurements) CUDA_TRY(cudaMemcpy(n_measurements + i, h->b.count + ids[i], sizeof(int32_t), cudaMemcpyDeviceToHost));
  }
  return VSLAM_OK;
}

int32_t vslam_landmark_map_size(const vslam_landmark_map* h) { return h ? h->n_landmarks : 0; }
int64_t vslam_landmark_map_launch_count(const vslam_landmark_map* h) { return h ? h->launches : 0; }

int vslam_landmark_optimizer_create(int32_t max_landmarks, int32_t max_measurements, int32_t max_frames, int device,
                                    vslam_landmark_optimizer** out) {
  if (!out) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  if (max_landmarks < 1 || max_measurements < 1 || max_frames < 1)
    return fail(VSLAM_ERR_INVALID_ARGUMENT, "capacities must be positive");
  int rc = require_device(device);
  if (rc) return rc;
  vslam_landmark_optimizer* h = new vslam_landmark_optimizer();
  h->device = device;
  h->max_landmarks = max_landmarks;
  h->max_measurements = max_measurements;
  h->max_frames = max_frames;
  bool ok = cudaSetDevice(device) == cudaSuccess && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
  auto dalloc = [&](void** p, size_t bytes) { ok = ok && cudaMalloc(p, bytes) == cudaSuccess; };
  dalloc((void**)&h->d_offsets, sizeof(int32_t) * ((size_t)max_landmarks + 1));
  dalloc((void**)&h->d_measurements, sizeof(LandmarkMeasurement) * (size_t)max_measurements);
  dalloc((void**)&h->d_poses, sizeof(double) * 24 * (size_t)max_frames);
  dalloc((void**)&h->d_world, sizeof(double) * 3 * (size_t)max_landmarks);
  dalloc((void**)&h->d_updates, sizeof(uint32_t) * (size_t)max_landmarks);
  dalloc((void**)&h->d_outcome, (size_t)max_landmarks);
  dalloc((void**)&h->d_iterations, sizeof(int32_t) * (size_t)max_landmarks);
  if (!ok) {
    const cudaError_t e = cudaGetLastError();
    vslam_landmark_optimizer_destroy(h);
    return fail(VSLAM_ERR_CUDA, "landmark optimizer allocation failed: %s", cudaGetErrorString(e));
  }
  *out = h;
  return VSLAM_OK;
}

int vslam_landmark_optimizer_destroy(vslam_landmark_optimizer* h) {
  if (!h) return VSLAM_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  cudaFree(h->d_offsets);
  cudaFree(h->d_measurements);
  cudaFree(h->d_poses);
  cudaFree(h->d_world);
  cudaFree(h->d_updates);
  cudaFree(h->d_outcome);
  cudaFree(h->d_iterations);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return VSLAM_OK;
}

int vslam_landmark_optimizer_update(vslam_landmark_optimizer* h, int32_t n_landmarks, const int32_t* measurement_offsets,
                                    const vslam_landmark_measurement* measurements, int32_t n_frames,
                                    const double* world_to_camera_left, const double* camera_left_to_world,
                                    uint32_t maximum_number_of_iterations, double maximum_error_squared_meters,
                                    double* world_coordinates, uint32_t* number_of_updates, uint8_t* outcome,
                                    int32_t* iterations) {
  VSLAM_NVTX("vslam_landmark_optimizer_update [PoseTracker3D::_updatePoints]");
  if (!h) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null handle");
  if (n_landmarks < 0 || n_landmarks > h->max_landmarks)
    return fail(VSLAM_ERR_CAPACITY, "%d landmarks exceed the capacity %d", n_landmarks, h->max_landmarks);
  if (n_landmarks == 0) return VSLAM_OK;
  if (!measurement_offsets || !measurements || !world_to_camera_left || !camera_left_to_world || !world_coordinates ||
      !number_of_updates)
    return fail(VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  if (n_frames < 1 || n_frames > h->max_frames)
    return fail(VSLAM_ERR_CAPACITY, "%d frames exceed the capacity %d", n_frames, h->max_frames);
  if (measurement_offsets[0] != 0) return fail(VSLAM_ERR_INVALID_ARGUMENT, "measurement_offsets[0] must be 0");
  for (int32_t i = 0; i < n_landmarks; ++i)
    if (measurement_offsets[i + 1] <= measurement_offsets[i])   // a landmark is born with >= 1 measurement (landmark.cpp:21-31)
      return fail(VSLAM_ERR_INVALID_ARGUMENT, "landmark %d has no measurement", i);
  const int32_t total = measurement_offsets[n_landmarks];
  if (total > h->max_measurements)
    return fail(VSLAM_ERR_CAPACITY, "%d measurements exceed the capacity %d", total, h->max_measurements);
  for (int32_t m = 0; m < total; ++m)
    if (measurements[m].frame < 0 || measurements[m].frame >= n_frames)
      return fail(VSLAM_ERR_INVALID_ARGUMENT, "measurement %d names frame %d of %d", m, measurements[m].frame, n_frames);
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  double* d_c2w = h->d_poses + 12 * (size_t)h->max_frames;
  CUDA_TRY(cudaMemcpyAsync(h->d_offsets, measurement_offsets, sizeof(int32_t) * ((size_t)n_landmarks + 1), cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(h->d_measurements, measurements, sizeof(LandmarkMeasurement) * (size_t)total, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(h->d_poses, world_to_camera_left, sizeof(double) * 12 * (size_t)n_frames, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(d_c2w, camera_left_to_world, sizeof(double) * 12 * (size_t)n_frames, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(h->d_world, world_coordinates, sizeof(double) * 3 * (size_t)n_landmarks, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(h->d_updates, number_of_updates, sizeof(uint32_t) * (size_t)n_landmarks, cudaMemcpyHostToDevice, s));
  launch_landmark_update(n_landmarks, h->d_offsets, h->d_measurements, h->d_poses, d_c2w, maximum_number_of_iterations,
                         maximum_error_squared_meters, h->d_world, h->d_updates, h->d_outcome, h->d_iterations, s);
  ++h->launches;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(world_coordinates, h->d_world, sizeof(double) * 3 * (size_t)n_landmarks, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(number_of_updates, h->d_updates, sizeof(uint32_t) * (size_t)n_landmarks, cudaMemcpyDeviceToHost, s));
  if (outcome) CUDA_TRY(cudaMemcpyAsync(outcome, h->d_outcome, (size_t)n_landmarks, cudaMemcpyDeviceToHost, s));
  if (iterations) CUDA_TRY(cudaMemcpyAsync(iterations, h->d_iterations, sizeof(int32_t) * (size_t)n_landmarks, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return VSLAM_OK;
}

int64_t vslam_landmark_optimizer_launch_count(const vslam_landmark_optimizer* h) { return h ? h->launches : 0; }

int32_t vslam_format_trajectory_kitti(const double robot_to_world[12], char* line, int32_t capacity) {
  return format_trajectory_kitti(robot_to_world, line, capacity);
}

int32_t vslam_format_trajectory_tum(double timestamp_seconds, const double robot_to_world[12], char* line, int32_t capacity) {
  return format_trajectory_tum(timestamp_seconds, robot_to_world, line, capacity);
}

int vslam_write_trajectory(const char* filename, int format, int32_t n_frames, const double* robot_to_world,
                           const double* timestamps_seconds) {
  if (!filename || !robot_to_world || n_frames < 0) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  if (format != VSLAM_TRAJECTORY_KITTI && format != VSLAM_TRAJECTORY_TUM)
    return fail(VSLAM_ERR_INVALID_ARGUMENT, "unknown trajectory format %d", format);
  if (format == VSLAM_TRAJECTORY_TUM && !timestamps_seconds)
    return fail(VSLAM_ERR_INVALID_ARGUMENT, "the TUM format needs timestamps");
  FILE* f = std::fopen(filename, "w");                       // overwriting, as the reference (world_map.cpp:196, :230)
  if (!f) return fail(VSLAM_ERR_INVALID_ARGUMENT, "cannot open %s", filename);
  char line[512];
  for (int32_t i = 0; i < n_frames; ++i) {
    const int n = format == VSLAM_TRAJECTORY_KITTI
                      ? format_trajectory_kitti(robot_to_world + 12 * (size_t)i, line, (int)sizeof(line))
                      : format_trajectory_tum(timestamps_seconds[i], robot_to_world + 12 * (size_t)i, line, (int)sizeof(line));
    if (n <= 0 || n >= (int)sizeof(line) || std::fwrite(line, 1, (size_t)n, f) != (size_t)n) {
      std::fclose(f);
      return fail(VSLAM_ERR_INVALID_ARGUMENT, "writing %s failed at frame %d", filename, i);
    }
  }
  std::fclose(f);
  return VSLAM_OK;
}

void vslam_solve3(const double A[9], const double rhs[3], double x[3]) { solve3(A, rhs, x); }

}  // extern "C"
