// landmark_api.cu -- C ABI of SURVEY.md 8f row 4 (include/vslam_b200.h): the batched landmark refinement
// (Landmark::update, reference src/types/landmark.cpp:66-152, driven by PoseTracker3D::_updatePoints,
// src/position_tracking/pose_tracker_3d.cpp:475-549) and the trajectory wire formats
// (WorldMap::writeTrajectoryKITTI / writeTrajectoryTUM, src/types/world_map.cpp:183-252).
#include <cuda_runtime.h>

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

extern "C" {

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
