// landmark.cu -- K13 landmark_update_kernel: Landmark::update (reference src/types/landmark.cpp:66-152), the per-landmark
// 3 x 3 Gauss-Newton refinement that PoseTracker3D::_updatePoints (src/position_tracking/pose_tracker_3d.cpp:475-549)
// runs for every tracked landmark of a frame.  Landmarks are independent: one WARP per landmark.
//
// The lanes evaluate the measurements of a chunk in parallel (transform, error, robust weight, the 9 + 3 + 1 terms of
// J^T W J, J^T W e and chi^2); the terms then meet in shared memory and lanes 0..12 add them IN MEASUREMENT ORDER, which
// is the reference's loop order: with -fmad=false and the oracle's expression order (oracle/c/vslam_oracle.c,
// orc_landmark_update) the refined positions are bit-identical to the CPU restatement, not merely close.
#include "gn_math.h"
#include "kernels.cuh"

namespace vslam {

namespace {

constexpr int kLmWarps = 4;      // landmarks per CTA
constexpr int kLmTerms = 13;     // H (9, the reference accumulates the full matrix), b (3), total error

// ordered sum of the lanes' terms: lane c < n_terms adds s[c][0 .. count-1] to its running sum
__device__ __forceinline__ void ordered_add(double (*s)[33], const double* terms, int n_terms, int count, int lane,
                                            double& running) {
  for (int c = 0; c < n_terms; ++c) s[c][lane] = terms[c];
  __syncwarp();
  if (lane < n_terms) {
    // the additions are a dependent chain (measurement order is the parity contract), the loads are not: eight at a time
    // are issued ahead of the chain instead of one shared-memory round trip per addition
    const double* row = s[lane];
    int l = 0;
    for (; l + 8 <= count; l += 8) {
      const double v0 = row[l], v1 = row[l + 1], v2 = row[l + 2], v3 = row[l + 3];
      const double v4 = row[l + 4], v5 = row[l + 5], v6 = row[l + 6], v7 = row[l + 7];
      running = running + v0;
      running = running + v1;
      running = running + v2;
      running = running + v3;
      running = running + v4;
      running = running + v5;
      running = running + v6;
      running = running + v7;
    }
    for (; l < count; ++l) running = running + row[l];
  }
  __syncwarp();
}

// solve3 of gn_math.h (Eigen::FullPivLU<Matrix3>::solve, landmark.cpp:136) with every index a compile-time constant:
// the 3 x 4 augmented system stays in registers (the generic form indexes its arrays with the run-time pivot position
// and lives in local memory: 160 bytes of stack per thread), the pivot search is a chain of selects in the scalar scan's
// order (strict `>`: the first maximum wins, a NaN never does), the row / column swaps are selects on the pivot position
// and the back substitution is straight-line code at full rank -- a branch region costs a warp ~100 cycles of latency on
// this machine, and an iteration of Landmark::update is one long dependency chain.  Same operations in
// the same order: bit-identical.
template <int K>
__device__ __forceinline__ void lu3_step(double (&a)[3][4], int (&perm)[3], int& rank, double& maxpivot) {
  if (K >= rank) return;                 // a zero pivot ended the elimination (the break of the scalar form)
  double biggest = -1;
  int pr = K, pc = K;
#pragma unroll
  for (int i = K; i < 3; ++i)
#pragma unroll
    for (int j = K; j < 3; ++j) {
      const double v = fabs(a[i][j]);
      const bool g = v > biggest;
      biggest = g ? v : biggest;
      pr = g ? i : pr;
      pc = g ? j : pc;
    }
  if (biggest == 0) {
    rank = K;
    return;
  }
  maxpivot = biggest > maxpivot ? biggest : maxpivot;
  // rows K and pr (all four columns: what lies left of K is never read again, the right-hand side follows)
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const double vk = a[K][j];
    double vp = vk;
#pragma unroll
    for (int r = K + 1; r < 3; ++r) vp = pr == r ? a[r][j] : vp;
#pragma unroll
    for (int r = K + 1; r < 3; ++r) a[r][j] = pr == r ? vk : a[r][j];
    a[K][j] = vp;
  }
  // columns K and pc (every row: the finished rows of U follow the permutation)
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double vk = a[i][K];
    double vp = vk;
#pragma unroll
    for (int c = K + 1; c < 3; ++c) vp = pc == c ? a[i][c] : vp;
#pragma unroll
    for (int c = K + 1; c < 3; ++c) a[i][c] = pc == c ? vk : a[i][c];
    a[i][K] = vp;
  }
  {
    const int vk = perm[K];
    int vp = vk;
#pragma unroll
    for (int c = K + 1; c < 3; ++c) vp = pc == c ? perm[c] : vp;
#pragma unroll
    for (int c = K + 1; c < 3; ++c) perm[c] = pc == c ? vk : perm[c];
    perm[K] = vp;
  }
#pragma unroll
  for (int i = K + 1; i < 3; ++i) {
    const double f = a[i][K] / a[K][K];
#pragma unroll
    for (int j = K + 1; j < 4; ++j) a[i][j] = a[i][j] - f * a[K][j];
  }
}

__device__ __forceinline__ void solve3_registers(const double* H, const double* rhs, double* x) {
  double a[3][4];
  int perm[3] = {0, 1, 2};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) a[i][j] = H[i * 3 + j];
    a[i][3] = rhs[i];
  }
  int rank = 3;
  double maxpivot = 0;
  lu3_step<0>(a, perm, rank, maxpivot);
  lu3_step<1>(a, perm, rank, maxpivot);
  lu3_step<2>(a, perm, rank, maxpivot);
  {
    int r = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) r += (i < rank && fabs(a[i][i]) > maxpivot * (2.220446049250313e-16 * 3)) ? 1 : 0;
    rank = r;
  }
  double y[3] = {0, 0, 0};
  if (rank == 3) {
    y[2] = a[2][3] / a[2][2];
    y[1] = (a[1][3] - a[1][2] * y[2]) / a[1][1];
    y[0] = ((a[0][3] - a[0][1] * y[1]) - a[0][2] * y[2]) / a[0][0];
  } else {
#pragma unroll
    for (int i = 2; i >= 0; --i)
      if (i < rank) {
        double sum = a[i][3];
#pragma unroll
        for (int j = 0; j < 3; ++j)
          if (j > i && j < rank) sum -= a[i][j] * y[j];
        y[i] = sum / a[i][i];
      }
  }
#pragma unroll
  for (int t = 0; t < 3; ++t) x[t] = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int t = 0; t < 3; ++t) x[t] = perm[i] == t ? y[i] : x[t];
}

// where the measurement history of a landmark lives: one CSR segment of a caller-provided array (vslam_landmark_optimizer)
// or the chained 32-entry blocks of the device-resident map (vslam_landmark_map)
struct CsrHistory {
  const LandmarkMeasurement* ms;
  __device__ __forceinline__ LandmarkMeasurement at(int m) const { return ms[m]; }
};
struct BlockHistory {
  const int32_t* table;                 // block ids of this landmark
  const LandmarkMeasurement* blocks;    // [n_blocks][32]
  __device__ __forceinline__ LandmarkMeasurement at(int m) const { return blocks[(size_t)table[m >> 5] * 32 + (m & 31)]; }
};

// Landmark::update for landmark `lm` (state in world / number_of_updates), history `ms` of n measurements; results of
// the call are also written at position `slot` of outcome / iterations (and of the compact copies when given)
template <class History>
__device__ __forceinline__ void landmark_update_warp(const History ms, int n, int lm, int slot, double (*s)[33],
                                                     const double* __restrict__ world_to_camera,
                                                     const double* __restrict__ camera_to_world, uint32_t max_iterations,
                                                     double max_err2, double* __restrict__ world,
                                                     uint32_t* __restrict__ number_of_updates, uint8_t* __restrict__ outcome,
                                                     int32_t* __restrict__ iterations, double* __restrict__ world_out,
                                                     uint32_t* __restrict__ updates_out) {
  const int lane = threadIdx.x & 31;
  double x[3] = {world[3 * (size_t)lm], world[3 * (size_t)lm + 1], world[3 * (size_t)lm + 2]};   // :82
  const uint32_t updates_so_far = number_of_updates[lm];
  double total_previous = 0;                                                                      // :88
  int result = 0;
  uint32_t it = 0;
  // the measurement and the pose of this lane in the FIRST chunk stay in registers across the iterations (most histories
  // are one chunk; an iteration otherwise starts with two dependent trips to memory: measurement -> its frame's pose)
  LandmarkMeasurement q_first = {};
  double w_first[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) w_first[i] = 0.0;
  if (lane < n) {
    q_first = ms.at(lane);
    const double* W = world_to_camera + 12 * (size_t)q_first.frame;
#pragma unroll
    for (int i = 0; i < 12; ++i) w_first[i] = __ldg(W + i);
  }
  for (; it < max_iterations; ++it) {                                                             // :91
    double running = 0;                                 // lane c: H[c] (c < 9), b[c - 9] (c < 12), total (c == 12)
    uint32_t outliers = 0;
    for (int c0 = 0; c0 < n; c0 += 32) {                                                          // :98
      const int m = c0 + lane;
      double terms[kLmTerms];
#pragma unroll
      for (int c = 0; c < kLmTerms; ++c) terms[c] = 0.0;   // +0.0 is neutral for sums that start at +0.0
      bool outlier = false;
      if (m < n) {
        LandmarkMeasurement q = q_first;
        double w_[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) w_[i] = w_first[i];
        if (c0 > 0) {
          q = ms.at(m);
          const double* W = world_to_camera + 12 * (size_t)q.frame;
#pragma unroll
          for (int i = 0; i < 12; ++i) w_[i] = __ldg(W + i);
        }
        double p[3], e[3];
#pragma unroll
        for (int r = 0; r < 3; ++r)                                                               // :102
          p[r] = ((w_[4 * r] * x[0] + w_[4 * r + 1] * x[1]) + w_[4 * r + 2] * x[2]) + w_[4 * r + 3];
        if (p[2] <= 0) {                                                                          // :103-106
          outlier = true;
        } else {
#pragma unroll
          for (int r = 0; r < 3; ++r) e[r] = p[r] - q.camera_coordinates[r];                      // :109
          double w = q.inverse_depth_meters;                                                      // :112
          const double err2 = ((e[0] * w) * e[0] + (e[1] * w) * e[1]) + (e[2] * w) * e[2];        // :115
          terms[12] = err2;                                                                       // :116
          if (err2 > max_err2) {                                                                  // :119-122
            w = w * (max_err2 / err2);
            outlier = true;
          }
#pragma unroll
          for (int i = 0; i < 3; ++i) {                                                           // :125-132
            const double jw0 = w_[i] * w, jw1 = w_[4 + i] * w, jw2 = w_[8 + i] * w;
#pragma unroll
            for (int j = 0; j < 3; ++j) terms[3 * i + j] = (jw0 * w_[j] + jw1 * w_[4 + j]) + jw2 * w_[8 + j];
            terms[9 + i] = (jw0 * e[0] + jw1 * e[1]) + jw2 * e[2];
          }
        }
      }
      outliers += __popc(__ballot_sync(0xffffffffu, outlier));
      ordered_add(s, terms, kLmTerms, min(32, n - c0), lane, running);
    }
    double H[9], nb[3], dx[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) H[i] = __shfl_sync(0xffffffffu, running, i);
#pragma unroll
    for (int i = 0; i < 3; ++i) nb[i] = -__shfl_sync(0xffffffffu, running, 9 + i);
    const double total = __shfl_sync(0xffffffffu, running, 12);
    solve3_registers(H, nb, dx);                                                                  // :136 (every lane)
#pragma unroll
    for (int i = 0; i < 3; ++i) x[i] = x[i] + dx[i];
    if (fabs(total - total_previous) < 1e-5 || it == 999) {                                       // :139
      const uint32_t inliers = (uint32_t)n - outliers;                                            // :140
      result = 3;
      if (inliers > updates_so_far) {                                                      // :143-147
        if (lane == 0) {
          world[3 * (size_t)lm] = x[0];
          world[3 * (size_t)lm + 1] = x[1];
          world[3 * (size_t)lm + 2] = x[2];
          number_of_updates[lm] = inliers;
        }
        result = 1;
      } else if (inliers < outliers) {                                                            // :150-160
        double acc = 0;
        for (int c0 = 0; c0 < n; c0 += 32) {
          const int m = c0 + lane;
          double terms[3] = {0.0, 0.0, 0.0};
          if (m < n) {
            const LandmarkMeasurement q = ms.at(m);
            const double* C = camera_to_world + 12 * (size_t)q.frame;
#pragma unroll
            for (int r = 0; r < 3; ++r)
              terms[r] = ((__ldg(C + 4 * r) * q.camera_coordinates[0] + __ldg(C + 4 * r + 1) * q.camera_coordinates[1]) +
                          __ldg(C + 4 * r + 2) * q.camera_coordinates[2]) + __ldg(C + 4 * r + 3);
          }
          ordered_add(s, terms, 3, min(32, n - c0), lane, acc);
        }
        if (lane < 3) world[3 * (size_t)lm + lane] = acc / n;
        result = 2;
      }
      ++it;
      break;
    }
    total_previous = total;                                                                       // :166
  }
  if (lane == 0) {
    if (outcome) outcome[slot] = (uint8_t)result;
    if (iterations) iterations[slot] = (int32_t)it;
  }
  if (world_out) {     // the state of the landmark after the call, compact (one device -> host copy for the frame)
    __syncwarp();
    if (lane < 3) world_out[3 * (size_t)slot + lane] = world[3 * (size_t)lm + lane];
    if (lane == 0) updates_out[slot] = number_of_updates[lm];
  }
}

__global__ void __launch_bounds__(kLmWarps * 32) landmark_update_kernel(
    int n_landmarks, const int32_t* __restrict__ offsets, const LandmarkMeasurement* __restrict__ measurements,
    const double* __restrict__ world_to_camera, const double* __restrict__ camera_to_world, uint32_t max_iterations,
    double max_err2, double* __restrict__ world, uint32_t* __restrict__ number_of_updates,
    uint8_t* __restrict__ outcome, int32_t* __restrict__ iterations) {
  __shared__ double s_terms[kLmWarps][kLmTerms][33];
  const int warp = threadIdx.x >> 5;
  const int lm = blockIdx.x * kLmWarps + warp;
  if (lm >= n_landmarks) return;                       // warps are independent (no block barrier)
  const int begin = offsets[lm], n = offsets[lm + 1] - begin;
  landmark_update_warp(CsrHistory{measurements + begin}, n, lm, lm, s_terms[warp], world_to_camera, camera_to_world,
                       max_iterations, max_err2, world, number_of_updates, outcome, iterations, nullptr, nullptr);
}

// ---- the device-resident landmark map (vslam_landmark_map): histories stay in HBM as chains of 32-measurement blocks;
// a frame appends ONE measurement per tracked landmark and refines it, nothing is re-uploaded.
// append: measurement number m of landmark lm goes to block table[lm][m / 32] (a new block is taken from the pool when
// m is a multiple of 32), slot m % 32.  Returns false when the pool or the table is exhausted.
__device__ __forceinline__ bool append_measurement(const LandmarkMapBuffers& b, int lm, const LandmarkMeasurement& q) {
  const int m = b.count[lm];
  int32_t* table = b.table + (size_t)lm * b.blocks_per_landmark;
  if ((m & 31) == 0) {
    if ((m >> 5) >= b.blocks_per_landmark) return false;
    const int blk = atomicAdd(b.next_block, 1);
    if (blk >= b.max_blocks) return false;
    table[m >> 5] = blk;
  }
  b.blocks[(size_t)table[m >> 5] * 32 + (m & 31)] = q;
  b.count[lm] = m + 1;
  return true;
}

// Landmark::Landmark (landmark.cpp:8-33) for n_new landmarks: the track of framepoints that gives birth to the landmark
// becomes its first measurements (newest first, as the constructor walks the track), _number_of_updates their number,
// _world_coordinates the average the host computed from the framepoints' world coordinates.  One thread per landmark.
__global__ void __launch_bounds__(128) landmark_create_kernel(LandmarkMapBuffers b, int n_new, int first_id,
                                                              const int32_t* __restrict__ track_offsets,
                                                              const LandmarkMeasurement* __restrict__ tracks,
                                                              const double* __restrict__ world_init) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_new) return;
  const int lm = first_id + i;
  b.count[lm] = 0;
  for (int k = track_offsets[i]; k < track_offsets[i + 1]; ++k)
    if (!append_measurement(b, lm, tracks[k])) atomicExch(b.error, 1);
  for (int d = 0; d < 3; ++d) b.world[3 * (size_t)lm + d] = world_init[3 * (size_t)i + d];
  b.updates[lm] = (uint32_t)(track_offsets[i + 1] - track_offsets[i]);
}

// PoseTracker3D::_updatePoints for the landmarks a frame tracked (pose_tracker_3d.cpp:486-519): Landmark::update =
// append the framepoint's measurement (landmark.cpp:71-79), then the Gauss-Newton refinement over the whole history.
__global__ void __launch_bounds__(kLmWarps * 32) landmark_map_update_kernel(
    LandmarkMapBuffers b, int n, const int32_t* __restrict__ ids, const double* __restrict__ camera_coordinates, int frame,
    uint32_t max_iterations, double max_err2, uint8_t* __restrict__ outcome, int32_t* __restrict__ iterations,
    double* __restrict__ world_out, uint32_t* __restrict__ updates_out) {
  __shared__ double s_terms[kLmWarps][kLmTerms][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * kLmWarps + warp;
  if (i >= n) return;
  const int lm = ids[i];
  int ok = 1;
  if (lane == 0) {
    LandmarkMeasurement q;                                       // Landmark::Measurement(framepoint), landmark.h:21-23
    q.frame = frame;
    q.reserved = 0;
    for (int d = 0; d < 3; ++d) q.camera_coordinates[d] = camera_coordinates[3 * (size_t)i + d];
    q.inverse_depth_meters = 1 / q.camera_coordinates[2];
    ok = append_measurement(b, lm, q) ? 1 : 0;
    if (!ok) atomicExch(b.error, 1);
    __threadfence_block();
  }
  ok = __shfl_sync(0xffffffffu, ok, 0);
  if (!ok) return;
  const int count = b.count[lm];
  landmark_update_warp(BlockHistory{b.table + (size_t)lm * b.blocks_per_landmark, b.blocks}, count, lm, i, s_terms[warp],
                       b.world_to_camera, b.camera_to_world, max_iterations, max_err2, b.world, b.updates, outcome,
                       iterations, world_out, updates_out);
}

}  // namespace

void launch_landmark_create(const LandmarkMapBuffers& b, int n_new, int first_id, const int32_t* track_offsets,
                            const LandmarkMeasurement* tracks, const double* world_init, cudaStream_t stream) {
  if (n_new <= 0) return;
  landmark_create_kernel<<<(n_new + 127) / 128, 128, 0, stream>>>(b, n_new, first_id, track_offsets, tracks, world_init);
}

void launch_landmark_map_update(const LandmarkMapBuffers& b, int n, const int32_t* ids, const double* camera_coordinates,
                                int frame, uint32_t max_iterations, double max_err2, uint8_t* outcome, int32_t* iterations,
                                double* world_out, uint32_t* updates_out, cudaStream_t stream) {
  if (n <= 0) return;
  landmark_map_update_kernel<<<(n + kLmWarps - 1) / kLmWarps, kLmWarps * 32, 0, stream>>>(
      b, n, ids, camera_coordinates, frame, max_iterations, max_err2, outcome, iterations, world_out, updates_out);
}

void launch_landmark_update(int n_landmarks, const int32_t* offsets, const LandmarkMeasurement* measurements,
                            const double* world_to_camera, const double* camera_to_world, uint32_t max_iterations,
                            double max_err2, double* world, uint32_t* number_of_updates, uint8_t* outcome,
                            int32_t* iterations, cudaStream_t stream) {
  if (n_landmarks <= 0) return;
  landmark_update_kernel<<<(n_landmarks + kLmWarps - 1) / kLmWarps, kLmWarps * 32, 0, stream>>>(
      n_landmarks, offsets, measurements, world_to_camera, camera_to_world, max_iterations, max_err2, world,
      number_of_updates, outcome, iterations);
}

}  // namespace vslam
