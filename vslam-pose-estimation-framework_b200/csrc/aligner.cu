// aligner.cu -- K7 stereouv / K8 uvd linearize: fused per-correspondence error + Jacobian + robust kernel and a
// deterministic grid-wide reduction of H (21 unique), b (6), total error and inlier count.
//
// Replaces StereoUVAligner::linearize (reference src/aligners/stereouv_aligner.cpp:72-187) and
// UVDAligner::linearize (src/aligners/uvd_aligner.cpp:77-171).  real = double (src/types/definitions.h:52).
//
// This file is compiled with -fmad=false: the per-point expressions below are written in the same order as the
// CPU oracle (oracle/c/vslam_oracle.c), so errors[] and the inlier classification are bit-identical to it; only
// the summation order of H/b differs (per-thread partials, warp shuffle tree, per-block partials, and ONE atomic
// ticket after which the last block adds the block partials in a fixed order -> run-to-run deterministic).
#include <cooperative_groups.h>

#include <cstdlib>

#include "gn_math.h"
#include "kernels.cuh"

namespace vslam {

namespace cg = cooperative_groups;

namespace {

constexpr int kAcc = 29;   // 21 H + 6 b + total_error + inliers
constexpr int kThreads = 256;

struct Pose {
  double T[12];
};

__device__ __forceinline__ int tri(int i, int j) {   // upper-triangle index, i <= j
  return i * 6 - (i * (i - 1)) / 2 + (j - i);
}

// One correspondence: error, robust kernel, Jacobian, and its contribution to acc[0..28].
// Returns the error (chi, or -1 when skipped) and the inlier flag exactly as linearize() leaves them in
// _errors[u] / _inliers[u].  fx[] = fixed measurement, om[] = information scalars (1 for StereoUV, 2 for UVD).
template <int KIND>
__device__ __forceinline__ void accumulate_point(const double m0, const double m1, const double m2, const double* fx,
                                                 const double* om, const double wt, const double* T,
                                                 const AlignerCamera& cam, const int ignore_outliers,
                                                 const double kernel, double (&acc)[kAcc], double& err, uint8_t& inl) {
  constexpr int D = KIND == 0 ? 4 : 3;
  const double* K = cam.K;
  err = -1.0;      // :82-84 / :88-90
  inl = 0;
  double pc[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) pc[i] = T[4 * i] * m0 + T[4 * i + 1] * m1 + T[4 * i + 2] * m2 + T[4 * i + 3];
  bool use = KIND == 0 ? !(pc[2] < cam.min_depth) : !(pc[2] <= cam.min_depth);   // :88 / :95
  double e[D], w[D], J[D][6];
  double abc[3], abr[3];
  if (use) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      abc[i] = K[3 * i] * pc[0] + K[3 * i + 1] * pc[1] + K[3 * i + 2] * pc[2];
      abr[i] = abc[i] + cam.baseline[i];
    }
    const double ul = abc[0] / abc[2], vl = abc[1] / abc[2];
    if (ul < 0 || ul > cam.cols || vl < 0 || vl > cam.rows) use = false;        // :103-106 / :109-112
    if (KIND == 0) {
      const double ur = abr[0] / abr[2], vr = abr[1] / abr[2];
      if (ur < 0 || ur > cam.cols || vr < 0 || vr > cam.rows) use = false;      // :107-110
      e[0] = ul - fx[0];
      e[1] = vl - fx[1];
      e[2] = ur - fx[2];
      e[3] = vr - fx[3];
#pragma unroll
      for (int d = 0; d < D; ++d) w[d] = om[0];
    } else {
      e[0] = ul - fx[0];
      e[1] = vl - fx[1];
      e[2] = pc[2] - fx[2];
      w[0] = w[1] = om[0];
      w[2] = om[1];
    }
  }
  if (use) {
    double chi = w[0] * e[0] * e[0];
#pragma unroll
    for (int d = 1; d < D; ++d) chi = chi + w[d] * e[d] * e[d];                  // :121 / :120
    err = chi;
    if (chi > kernel) {                                                          // :127-137 / :126-135
      if (ignore_outliers) {
        use = false;
      } else {
        const double s = kernel / chi;
#pragma unroll
        for (int d = 0; d < D; ++d) w[d] = w[d] * s;
      }
    } else {
      inl = 1;
      acc[28] += 1.0;
    }
  }
  if (!use) return;
  acc[27] += err;                                                                // :140 / :138

  // From here on only H and b are fed (tolerance 1e-10 relative: the reduction order differs from the reference's loop
  // anyway): the products are contracted with explicit fma() -- the file is compiled with -fmad=false so that
  // everything ABOVE (errors, inlier decisions, robust weights) keeps the oracle's uncontracted evaluation and stays
  // bit-identical.  ~140 fewer FP64 instructions per correspondence; the kernels are bound by the FP64 pipe.
  // K * [wt*I3 | -2*skew(p)]  (:143-152 / :145-161); zero terms of the dense product are dropped (exact)
  double kj[3][6];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    kj[i][0] = K[3 * i] * wt;
    kj[i][1] = K[3 * i + 1] * wt;
    kj[i][2] = K[3 * i + 2] * wt;
    kj[i][3] = fma(K[3 * i + 1], -2 * pc[2], K[3 * i + 2] * (2 * pc[1]));
    kj[i][4] = fma(K[3 * i], 2 * pc[2], K[3 * i + 2] * (-2 * pc[0]));
    kj[i][5] = fma(K[3 * i], -2 * pc[1], K[3 * i + 1] * (2 * pc[0]));
  }
  if (KIND == 0) {
    const double il = 1 / abc[2], ir = 1 / abr[2];                               // :155-158
    const double il2 = il * il, ir2 = ir * ir;
    const double al0 = -abc[0] * il2, al1 = -abc[1] * il2, ar0 = -abr[0] * ir2, ar1 = -abr[1] * ir2;
#pragma unroll
    for (int j = 0; j < 6; ++j) {                                                // :161-177
      J[0][j] = fma(il, kj[0][j], al0 * kj[2][j]);
      J[1][j] = fma(il, kj[1][j], al1 * kj[2][j]);
      J[2][j] = fma(ir, kj[0][j], ar0 * kj[2][j]);
      J[3][j] = fma(ir, kj[1][j], ar1 * kj[2][j]);
    }
  } else {
    const double iz = 1 / pc[2], iz2 = iz * iz;                                  // :141-142
    const double a0 = -abc[0] * iz2, a1 = -abc[1] * iz2;
#pragma unroll
    for (int j = 0; j < 6; ++j) {                                                // :155-161
      J[0][j] = fma(iz, kj[0][j], a0 * kj[2][j]);
      J[1][j] = fma(iz, kj[1][j], a1 * kj[2][j]);
      J[2][j] = kj[2][j];
    }
  }
  // H += J^T W J (upper triangle), b += J^T W e                                  (:183-184 / :167-168)
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double jw[D];
#pragma unroll
    for (int d = 0; d < D; ++d) jw[d] = J[d][i] * w[d];
#pragma unroll
    for (int j = i; j < 6; ++j) {
      double a = acc[tri(i, j)];
#pragma unroll
      for (int d = 0; d < D; ++d) a = fma(jw[d], J[d][j], a);
      acc[tri(i, j)] = a;
    }
    double a = acc[21 + i];
#pragma unroll
    for (int d = 0; d < D; ++d) a = fma(jw[d], e[d], a);
    acc[21 + i] = a;
  }
}

// warp shuffle tree + cross-warp sum; the block's kAcc totals end up in threads 0..kAcc-1 (return value)
__device__ __forceinline__ double block_reduce(const double (&acc)[kAcc], double (*s_part)[kAcc]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < kAcc; ++i) {
    double v = acc[i];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_part[warp][i] = v;
  }
  __syncthreads();
  double v = 0;
  if (threadIdx.x < kAcc)
    for (int w = 0; w < kThreads / 32; ++w) v += s_part[w][threadIdx.x];
  return v;
}

template <int KIND>
__global__ void __launch_bounds__(kThreads) linearize_kernel(int n, AlignerBuffers b, AlignerCamera cam, Pose pose,
                                                             int ignore_outliers, double kernel) {
  constexpr int D = KIND == 0 ? 4 : 3;
  constexpr int W = KIND == 0 ? 1 : 2;
  double acc[kAcc];
#pragma unroll
  for (int i = 0; i < kAcc; ++i) acc[i] = 0.0;

  for (int u = blockIdx.x * kThreads + threadIdx.x; u < n; u += gridDim.x * kThreads) {
    double fx[D], om[W], err;
    uint8_t inl;
#pragma unroll
    for (int d = 0; d < D; ++d) fx[d] = b.fixed[d * b.stride + u];
#pragma unroll
    for (int d = 0; d < W; ++d) om[d] = b.omega[d * b.stride + u];
    accumulate_point<KIND>(b.moving[u], b.moving[b.stride + u], b.moving[2 * b.stride + u], fx, om, b.wt[u], pose.T, cam,
                           ignore_outliers, kernel, acc, err, inl);
    b.errors[u] = err;
    b.inliers[u] = inl;
  }

  // ---- warp shuffle tree, then per-block partials
  __shared__ double s_part[kThreads / 32][kAcc];
  __shared__ bool s_last;
  const double total = block_reduce(acc, s_part);
  if (threadIdx.x < kAcc) b.partials[(size_t)blockIdx.x * 32 + threadIdx.x] = total;
  // ---- the one atomic stage: a ticket; the last block to arrive reduces the block partials in block order
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicInc(b.ticket, gridDim.x - 1);   // wraps to 0 after the last block: self-resetting
    s_last = t == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  {
    const int j = threadIdx.x & 31, part = threadIdx.x >> 5;   // 8 interleaved slices of the block list per value
    double v = 0;
    if (j < kAcc)
      for (int blk = part; blk < (int)gridDim.x; blk += kThreads / 32) v += __ldcg(&b.partials[(size_t)blk * 32 + j]);
    __syncthreads();
    if (j < kAcc) s_part[part][j] = v;
    __syncthreads();
    if (threadIdx.x < kAcc) {
      double s = 0;
      for (int w = 0; w < kThreads / 32; ++w) s += s_part[w][threadIdx.x];
      b.system[threadIdx.x] = s;
    }
  }
}

// solve6 of gn_math.h (Eigen::FullPivLU<Matrix6>::solve) by ONE WARP, bit-identical to the single-thread form: lane j < 6
// holds column j of the matrix, lane 6 the right-hand side as a seventh column (row swaps and eliminations act on it
// exactly as on the others).  The single-thread form indexes its arrays with the run-time pivot position, i.e. lives in
// local memory, and costs ~8 us per Gauss-Newton round; here every index is a compile-time constant or a select.
// Every lane of the warp must call; every lane receives x[6].
__device__ __forceinline__ void solve6_warp(const double* H /* 36, row-major */, const double* rhs, double x[6]) {
  const int lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  double a[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) a[r] = lane < 6 ? H[r * 6 + lane] : (lane == 6 ? rhs[r] : 0.0);
  int perm = lane;                 // lane j < 6: perm[j]
  int rank = 6;
  double maxpivot = 0;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    if (k < rank) {               // (uniform: after a zero pivot the remaining steps are skipped, as the break does)
      // ---- complete pivoting: the first element in row-major order that attains the maximum of |A[i][j]|, i, j >= k
      double best = -1;
      int bi = k;
      if (lane >= k && lane < 6) {
#pragma unroll
        for (int r = 0; r < 6; ++r)
          if (r >= k && fabs(a[r]) > best) {
            best = fabs(a[r]);
            bi = r;
          }
      }
      int bj = lane;
#pragma unroll
      for (int o = 4; o; o >>= 1) {
        const double ov = __shfl_xor_sync(full, best, o);
        const int oi = __shfl_xor_sync(full, bi, o), oj = __shfl_xor_sync(full, bj, o);
        if (ov > best || (ov == best && (oi < bi || (oi == bi && oj < bj)))) {
          best = ov;
          bi = oi;
          bj = oj;
        }
      }
      // lanes 0..7 now agree; everyone takes lane 0's
      const double biggest = __shfl_sync(full, best, 0);
      const int pr = __shfl_sync(full, bi, 0), pc = __shfl_sync(full, bj, 0);
      if (biggest == 0) {
        rank = k;
      } else {
        if (biggest > maxpivot) maxpivot = biggest;
        if (pr != k) {            // row swap (the right-hand side in lane 6 follows)
          const double vk = a[k];
          double vp = vk;
#pragma unroll
          for (int r = 0; r < 6; ++r)
            if (r == pr) vp = a[r];
#pragma unroll
          for (int r = 0; r < 6; ++r)
            if (r == pr) a[r] = vk;
          a[k] = vp;
        }
        {                          // column swap: lanes k and pc exchange their columns and their perm entry
          const int partner = pc != k ? (lane == k ? pc : (lane == pc ? k : lane)) : lane;
#pragma unroll
          for (int r = 0; r < 6; ++r) a[r] = __shfl_sync(full, a[r], partner);
          perm = __shfl_sync(full, perm, partner);
        }
        // ---- elimination: f_i = A[i][k] / A[k][k] from column k (lane k), then A[i][j] -= f_i A[k][j] in every column
        double f[6];
        const double pivot = __shfl_sync(full, a[k], k);
#pragma unroll
        for (int r = 0; r < 6; ++r) f[r] = r > k ? __shfl_sync(full, a[r], k) / pivot : 0.0;
        if (lane == k) {
#pragma unroll
          for (int r = 0; r < 6; ++r)
            if (r > k) a[r] = f[r];
        } else if (lane > k && lane <= 6) {
#pragma unroll
          for (int r = 0; r < 6; ++r)
            if (r > k) a[r] = a[r] - f[r] * a[k];
        }
      }
    }
  }
  // Eigen::FullPivLU::rank(): only pivots above |largest pivot| * epsilon * size are used by solve()
  {
    double diag = 0;
#pragma unroll
    for (int r = 0; r < 6; ++r)
      if (r == lane) diag = a[r];
    const bool used = lane < rank && fabs(diag) > maxpivot * (2.220446049250313e-16 * 6);
    rank = __popc(__ballot_sync(full, used));
  }
  // ---- back substitution, every lane redundantly on a gathered copy of U and the transformed right-hand side
  double U[6][6], bb[6], y[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    bb[i] = __shfl_sync(full, a[i], 6);
#pragma unroll
    for (int j = 0; j < 6; ++j)
      U[i][j] = j >= i ? __shfl_sync(full, a[i], j) : 0.0;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) y[i] = 0;
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    if (i < rank) {
      double sum = bb[i];
#pragma unroll
      for (int j = 0; j < 6; ++j)
        if (j > i && j < rank) sum = sum - U[i][j] * y[j];
      y[i] = sum / U[i][i];
    }
  }
  // x[perm[i]] = y[i]
#pragma unroll
  for (int t = 0; t < 6; ++t) x[t] = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int pi = __shfl_sync(full, perm, i);
#pragma unroll
    for (int t = 0; t < 6; ++t)
      if (t == pi) x[t] = y[i];
  }
}

// oneRound (:190-207 / :174-191) + the converge state machine (:213-247 / :197-233) for one finished linearisation:
// damped system, full-pivot solve, pose update, bookkeeping in *ctl (global memory in the grid kernel, shared memory in
// the cluster kernel).  Called by the WHOLE first warp of the block (the solve is warp-cooperative); lane 0 writes.
// `s_H` is a 36-double scratch in shared memory.
__device__ __forceinline__ void gn_step(GnControl* ctl, const double* s_sys, const double* s_T, double* s_H, int n,
                                        const GnParams& p) {
  const int lane = threadIdx.x & 31;
  {   // H as the single-thread form builds it: both triangles from the packed upper one, then the damping on the diagonal
    int k = 0;
    for (int i = 0; i < 6; ++i)
      for (int j = i; j < 6; ++j, ++k)
        if (lane == 0) s_H[i * 6 + j] = s_H[j * 6 + i] = s_sys[k];
    __syncwarp();
    if (lane < 6) s_H[lane * 6 + lane] += p.damping * n;
    __syncwarp();
  }
  double nb[6], dx[6];
  for (int i = 0; i < 6; ++i) nb[i] = -s_sys[21 + i];
  solve6_warp(s_H, nb, dx);
  if (lane != 0) return;
  double T[12];
  for (int i = 0; i < 12; ++i) T[i] = s_T[i];
  apply_update(dx, T);
  for (int i = 0; i < 12; ++i) ctl->T[i] = T[i];
  for (int i = 0; i < 36; ++i) ctl->H[i] = s_H[i];
  const double total_error = s_sys[27];
  const int inliers = (int)llrint(s_sys[28]);
  const int outliers = n - inliers;
  const double prev = ctl->total_error_previous;
  int rounds = ctl->rounds + 1, done = 0, converged = 0, phase = ctl->phase, it = ctl->iteration;
  if (phase == 0) {
    if (p.error_delta > fabs(prev - total_error)) {
      if (inliers > p.inlier_gate && inliers > outliers && p.max_iterations > 0) {
        phase = 1;      // inlier-only rounds (:224-236)
        it = 0;
      } else {
        done = converged = 1;
      }
    } else if (++it >= p.max_iterations) {
      done = 1;         // "system did not converge" (:250-255)
    }
  } else {
    if (fabs(prev - total_error) < p.error_delta || ++it >= p.max_iterations) done = converged = 1;
  }
  ctl->total_error_previous = total_error;
  ctl->rounds = rounds;
  ctl->phase = phase;
  ctl->iteration = it;
  ctl->ignore = phase;
  ctl->converged = converged;
  __threadfence();
  ctl->done = done;
}

// Fused Gauss-Newton: BaseAligner::converge (reference stereouv_aligner.cpp:210-264, uvd_aligner.cpp:194-248) as ONE
// persistent cooperative kernel -- per round: linearize (same per-point code and the same ordered reduction as
// linearize_kernel, so H and b are bit-identical to the stepwise path), grid barrier, block 0 solves the damped 6x6
// system, updates the pose (v2t, re-orthonormalisation) and advances the convergence state machine, grid barrier.
// No host round trip per round; errors[] / inliers[] hold the last round's values as in the reference.
template <int KIND>
__global__ void __launch_bounds__(kThreads) converge_kernel(int n, AlignerBuffers b, AlignerCamera cam, GnParams p,
                                                            GnControl* __restrict__ ctl) {
  constexpr int D = KIND == 0 ? 4 : 3;
  constexpr int W = KIND == 0 ? 1 : 2;
  cg::grid_group grid = cg::this_grid();
  __shared__ double s_part[kThreads / 32][kAcc];
  __shared__ double s_T[12];
  __shared__ double s_sys[32];
  __shared__ double s_H[36];
  __shared__ int s_ignore;

  for (;;) {
    if (threadIdx.x < 12) s_T[threadIdx.x] = __ldcg(&ctl->T[threadIdx.x]);
    if (threadIdx.x == 0) s_ignore = __ldcg(&ctl->ignore);
    __syncthreads();
    const int ignore_outliers = s_ignore;

    double acc[kAcc];
#pragma unroll
    for (int i = 0; i < kAcc; ++i) acc[i] = 0.0;
    for (int u = blockIdx.x * kThreads + threadIdx.x; u < n; u += gridDim.x * kThreads) {
      double fx[D], om[W], err;
      uint8_t inl;
#pragma unroll
      for (int d = 0; d < D; ++d) fx[d] = b.fixed[d * b.stride + u];
#pragma unroll
      for (int d = 0; d < W; ++d) om[d] = b.omega[d * b.stride + u];
      accumulate_point<KIND>(b.moving[u], b.moving[b.stride + u], b.moving[2 * b.stride + u], fx, om, b.wt[u], s_T, cam,
                             ignore_outliers, p.kernel, acc, err, inl);
      b.errors[u] = err;
      b.inliers[u] = inl;
    }
    const double total = block_reduce(acc, s_part);
    if (threadIdx.x < kAcc) b.partials[(size_t)blockIdx.x * 32 + threadIdx.x] = total;
    grid.sync();

    if (blockIdx.x == 0) {
      {  // the block partials in block order, 8 interleaved slices per value -- as in linearize_kernel
        const int j = threadIdx.x & 31, part = threadIdx.x >> 5;
        double v = 0;
        if (j < kAcc)
          for (int blk = part; blk < (int)gridDim.x; blk += kThreads / 32) v += __ldcg(&b.partials[(size_t)blk * 32 + j]);
        __syncthreads();
        if (j < kAcc) s_part[part][j] = v;
        __syncthreads();
        if (threadIdx.x < kAcc) {
          double s = 0;
          for (int w = 0; w < kThreads / 32; ++w) s += s_part[w][threadIdx.x];
          s_sys[threadIdx.x] = s;
          b.system[threadIdx.x] = s;
        }
        __syncthreads();
      }
      if (threadIdx.x < 32) gn_step(ctl, s_sys, s_T, s_H, n, p);
    }
    grid.sync();
    if (__ldcg(&ctl->done)) break;
  }
}

// The same loop for the problem sizes of a tracked frame (n <= 8 x 256 correspondences): ONE thread-block cluster
// instead of a cooperative grid.  The CTAs of the cluster are the blocks of the grid version -- same per-point code, same
// per-block partials, same order of the final sum, so pose and round count stay bit-identical -- but the two barriers of
// a round are cluster barriers (hardware, no round trip through global memory), the block partials and the control
// block live in shared memory and are read through distributed shared memory, and every thread keeps its ONE
// correspondence in registers across the rounds.  Launched as a plain kernel with a cluster dimension.
template <int KIND>
__global__ void __launch_bounds__(kThreads, 1) converge_cluster_kernel(int n, AlignerBuffers b, AlignerCamera cam, GnParams p,
                                                                    GnControl* __restrict__ ctl,
                                                                    const int32_t* __restrict__ n_device) {
  constexpr int D = KIND == 0 ? 4 : 3;
  constexpr int W = KIND == 0 ? 1 : 2;
  // fused frame (captured graph): the correspondence count is what track() left in device memory; nothing to align
  // without tracks (pose_tracker_3d.cpp:355: the tracker only optimises a frame that has points)
  if (n_device) {
    n = *n_device;
    if (n <= 0 || n > (int)(gridDim.x * kThreads)) return;
  }
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), n_blocks = (int)cluster.num_blocks();
  __shared__ double s_part[kThreads / 32][kAcc];
  __shared__ double s_total[32];        // this block's partial sums, read by block 0
  __shared__ double s_T[12];
  __shared__ double s_sys[32];
  __shared__ double s_H[36];
  __shared__ GnControl s_ctl;           // authoritative copy in block 0
  GnControl* ctl0 = cluster.map_shared_rank(&s_ctl, 0);

  if (rank == 0) {
    if (threadIdx.x < 12) s_ctl.T[threadIdx.x] = ctl->T[threadIdx.x];
    if (threadIdx.x == 0) {
      s_ctl.total_error_previous = ctl->total_error_previous;
      s_ctl.rounds = ctl->rounds;
      s_ctl.phase = ctl->phase;
      s_ctl.iteration = ctl->iteration;
      s_ctl.ignore = ctl->ignore;
      s_ctl.converged = ctl->converged;
      s_ctl.done = ctl->done;
    }
  }
  // this thread's correspondence (the grid version's block `rank`, thread threadIdx.x, first and only iteration)
  const int u = rank * kThreads + threadIdx.x;
  const bool mine = u < n;
  double m[3] = {0, 0, 0}, fx[D], om[W], wt = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) fx[d] = 0;
#pragma unroll
  for (int d = 0; d < W; ++d) om[d] = 0;
  if (mine) {
#pragma unroll
    for (int d = 0; d < 3; ++d) m[d] = b.moving[d * b.stride + u];
#pragma unroll
    for (int d = 0; d < D; ++d) fx[d] = b.fixed[d * b.stride + u];
#pragma unroll
    for (int d = 0; d < W; ++d) om[d] = b.omega[d * b.stride + u];
    wt = b.wt[u];
  }
  cluster.sync();

  double err = -1.0;
  uint8_t inl = 0;
  for (;;) {
    if (threadIdx.x < 12) s_T[threadIdx.x] = ctl0->T[threadIdx.x];
    const int ignore_outliers = ctl0->ignore;
    __syncthreads();
    double acc[kAcc];
#pragma unroll
    for (int i = 0; i < kAcc; ++i) acc[i] = 0.0;
    if (mine) accumulate_point<KIND>(m[0], m[1], m[2], fx, om, wt, s_T, cam, ignore_outliers, p.kernel, acc, err, inl);
    const double total = block_reduce(acc, s_part);
    if (threadIdx.x < kAcc) s_total[threadIdx.x] = total;
    cluster.sync();                     // every block's partial is in its shared memory

    if (rank == 0) {
      {  // the block partials in block order, 8 interleaved slices per value -- as in linearize_kernel (blocks without a
         // correspondence contribute +0.0, so a cluster larger than the grid of the stepwise path adds up to the same bits)
        const int j = threadIdx.x & 31, part = threadIdx.x >> 5;
        double v = 0;
        if (j < kAcc)
          for (int blk = part; blk < n_blocks; blk += kThreads / 32) v += cluster.map_shared_rank(s_total, blk)[j];
        __syncthreads();
        if (j < kAcc) s_part[part][j] = v;
        __syncthreads();
        if (threadIdx.x < kAcc) {
          double s = 0;
          for (int w = 0; w < kThreads / 32; ++w) s += s_part[w][threadIdx.x];
          s_sys[threadIdx.x] = s;
        }
        __syncthreads();
      }
      if (threadIdx.x < 32) gn_step(&s_ctl, s_sys, s_T, s_H, n, p);
    }
    cluster.sync();                     // block 0's control block is final for this round
    if (ctl0->done) break;
  }
  // errors[] / inliers[] hold the last round's values as in the reference; the system and the control block go to
  // global memory once
  if (mine) {
    b.errors[u] = err;
    b.inliers[u] = inl;
  }
  if (rank == 0) {
    if (threadIdx.x < kAcc) b.system[threadIdx.x] = s_sys[threadIdx.x];
    if (threadIdx.x < 12) ctl->T[threadIdx.x] = s_ctl.T[threadIdx.x];
    if (threadIdx.x < 36) ctl->H[threadIdx.x] = s_ctl.H[threadIdx.x];
    if (threadIdx.x == 0) {
      ctl->total_error_previous = s_ctl.total_error_previous;
      ctl->rounds = s_ctl.rounds;
      ctl->phase = s_ctl.phase;
      ctl->iteration = s_ctl.iteration;
      ctl->ignore = s_ctl.ignore;
      ctl->converged = s_ctl.converged;
      ctl->done = s_ctl.done;
    }
  }
  cluster.sync();                       // no block may exit while block 0 still reads its shared memory
}

// Batched form for independent stereo pairs: one WARP per pair linearises the StereoUV problem that aligns the
// pair's new framepoints against themselves -- StereoUVAligner::initialize (:10-69) fused in: _moving =
// cameraCoordinatesLeft, _fixed = (uL, vL, uR, vR), information = I4 (no landmark), w_t = min(max_depth/depth, 1).
// A pair holds a few hundred points and every point is a long dependent FP64 chain (three divisions): one CTA per
// pair with one point per thread and iteration keeps the chain short; the warps' partial sums meet in shared memory
// and are added in a fixed order (run-to-run deterministic).
constexpr int kPairWarps = 4;

__global__ void __launch_bounds__(kPairWarps * 32) linearize_pairs_kernel(const FramePointRecord* __restrict__ records,
                                                                          int record_stride, int n_pairs,
                                                                          const int32_t* __restrict__ n_out,
                                                                          AlignerCamera cam, Pose pose,
                                                                          int ignore_outliers, double kernel,
                                                                          double max_reliable_depth,
                                                                          int inverse_depth_weight,
                                                                          double* __restrict__ systems,
                                                                          double* __restrict__ errors,
                                                                          uint8_t* __restrict__ inliers) {
  __shared__ double s_part[kPairWarps][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pair = blockIdx.x;
  const int n = n_out[2 * pair];
  const FramePointRecord* rec = records + (size_t)pair * record_stride;
  double acc[kAcc];
#pragma unroll
  for (int i = 0; i < kAcc; ++i) acc[i] = 0.0;
  for (int u = threadIdx.x; u < n; u += kPairWarps * 32) {
    const FramePointRecord r = rec[u];
    double err = -1.0;
    uint8_t inl = 0;
    if (r.index_left >= 0) {
      const double fx[4] = {(double)r.xl, (double)r.yl, (double)r.xr, (double)r.yr};
      const double om[1] = {1.0};
      const double wt = inverse_depth_weight ? fmin(max_reliable_depth / r.camera[2], 1.0) : 1.0;   // :59-63
      accumulate_point<0>(r.camera[0], r.camera[1], r.camera[2], fx, om, wt, pose.T, cam, ignore_outliers, kernel, acc,
                          err, inl);
    }
    errors[(size_t)pair * record_stride + u] = err;
    inliers[(size_t)pair * record_stride + u] = inl;
  }
  double mine = 0.0;   // lane i ends up with the warp's total i
#pragma unroll
  for (int i = 0; i < kAcc; ++i) {
    double v = acc[i];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == i) mine = v;
  }
  s_part[warp][lane] = mine;
  __syncthreads();
  if (warp == 0 && lane < kAcc) {
    double v = s_part[0][lane];
#pragma unroll
    for (int w = 1; w < kPairWarps; ++w) v += s_part[w][lane];
    systems[(size_t)pair * 32 + lane] = v;
  }
}

}  // namespace

// grid of both the stepwise and the fused kernel: every block co-resident (the fused kernel is cooperative), and the
// SAME grid for both so that their ordered reductions -- hence H, b and the pose sequence -- are bit-identical
int aligner_grid(int n, int resident_blocks) {
  const int blocks = (n + kThreads - 1) / kThreads;
  return blocks < 1 ? 1 : (blocks < resident_blocks ? blocks : resident_blocks);
}

void launch_linearize(int kind, int n, const AlignerBuffers& b, const AlignerCamera& cam, const double T[12],
                      int ignore_outliers, double kernel, int grid, cudaStream_t stream) {
  Pose pose;
  for (int i = 0; i < 12; ++i) pose.T[i] = T[i];
  if (kind == 0) linearize_kernel<0><<<grid, kThreads, 0, stream>>>(n, b, cam, pose, ignore_outliers, kernel);
  else linearize_kernel<1><<<grid, kThreads, 0, stream>>>(n, b, cam, pose, ignore_outliers, kernel);
}

int converge_max_blocks_per_sm(int kind) {
  int n = 0;
  if (kind == 0) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, converge_kernel<0>, kThreads, 0);
  else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, converge_kernel<1>, kThreads, 0);
  return n;
}

cudaError_t launch_converge(int kind, int n, const AlignerBuffers& b, const AlignerCamera& cam, const GnParams& p,
                            GnControl* ctl, int grid, cudaStream_t stream) {
  int n_arg = n;
  AlignerBuffers b_arg = b;
  AlignerCamera cam_arg = cam;
  GnParams p_arg = p;
  void* args[] = {&n_arg, &b_arg, &cam_arg, &p_arg, &ctl};
  // the sizes of a tracked frame: one thread-block cluster (the same blocks, cheaper barriers); anything larger, or a
  // cluster size the device refuses: the cooperative grid
  static bool cluster_ok = std::getenv("VSLAM_NO_CLUSTER_CONVERGE") == nullptr;
  if (cluster_ok && grid <= 8 && n <= grid * kThreads) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = grid;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    const int32_t* no_device_count = nullptr;
    const cudaError_t e =
        kind == 0 ? cudaLaunchKernelEx(&cfg, converge_cluster_kernel<0>, n_arg, b_arg, cam_arg, p_arg, ctl, no_device_count)
                  : cudaLaunchKernelEx(&cfg, converge_cluster_kernel<1>, n_arg, b_arg, cam_arg, p_arg, ctl, no_device_count);
    if (e == cudaSuccess) return e;
    cudaGetLastError();                 // e.g. a cluster size this device does not schedule: use the grid kernel
  }
  const void* fn = kind == 0 ? (const void*)converge_kernel<0> : (const void*)converge_kernel<1>;
  return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kThreads), args, 0, stream);
}

namespace {
void cluster_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, int blocks, cudaStream_t stream) {
  *cfg = cudaLaunchConfig_t{};
  cfg->gridDim = dim3(blocks);
  cfg->blockDim = dim3(kThreads);
  cfg->stream = stream;
  attr->id = cudaLaunchAttributeClusterDimension;
  attr->val.clusterDim.x = blocks;
  attr->val.clusterDim.y = 1;
  attr->val.clusterDim.z = 1;
  cfg->attrs = attr;
  cfg->numAttrs = 1;
}
}  // namespace

int frame_step_cluster_blocks() {
  // 16 CTAs per cluster need the non-portable opt-in; the occupancy query says whether this device co-schedules them
  const bool opt_in = cudaFuncSetAttribute(converge_cluster_kernel<0>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
  for (int blocks : {16, 8}) {
    if (blocks > 8 && !opt_in) continue;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr;
    cluster_config(&cfg, &attr, blocks, nullptr);
    int clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&clusters, converge_cluster_kernel<0>, &cfg) == cudaSuccess && clusters > 0) return blocks;
    cudaGetLastError();
  }
  return 0;
}

cudaError_t launch_converge_frame(const AlignerBuffers& b, const AlignerCamera& cam, const GnParams& p, GnControl* ctl,
                                  const int32_t* n_device, int cluster_blocks, cudaStream_t stream) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr;
  cluster_config(&cfg, &attr, cluster_blocks, stream);
  return cudaLaunchKernelEx(&cfg, converge_cluster_kernel<0>, 0, b, cam, p, ctl, n_device);
}

void launch_linearize_pairs(const FramePointRecord* records, int record_stride, const int32_t* n_out, int n_pairs,
                            const AlignerCamera& cam, const double T[12], int ignore_outliers, double kernel,
                            double max_reliable_depth, int inverse_depth_weight, double* systems, double* errors,
                            uint8_t* inliers, cudaStream_t stream) {
  Pose pose;
  for (int i = 0; i < 12; ++i) pose.T[i] = T[i];
  linearize_pairs_kernel<<<n_pairs, kPairWarps * 32, 0, stream>>>(
      records, record_stride, n_pairs, n_out, cam, pose, ignore_outliers, kernel, max_reliable_depth,
      inverse_depth_weight, systems, errors, inliers);
}

}  // namespace vslam
