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
#include "kernels.cuh"

namespace vslam {

namespace {

constexpr int kAcc = 29;   // 21 H + 6 b + total_error + inliers
constexpr int kThreads = 256;

struct Pose {
  double T[12];
};

__device__ __forceinline__ int tri(int i, int j) {   // upper-triangle index, i <= j
  return i * 6 - (i * (i - 1)) / 2 + (j - i);
}

template <int KIND>
__global__ void __launch_bounds__(kThreads) linearize_kernel(int n, AlignerBuffers b, AlignerCamera cam, Pose pose,
                                                             int ignore_outliers, double kernel) {
  constexpr int D = KIND == 0 ? 4 : 3;
  double acc[kAcc];
#pragma unroll
  for (int i = 0; i < kAcc; ++i) acc[i] = 0.0;
  const double* K = cam.K;
  const double* T = pose.T;

  for (int u = blockIdx.x * kThreads + threadIdx.x; u < n; u += gridDim.x * kThreads) {
    double err = -1.0;      // :82-84 / :88-90
    uint8_t inl = 0;
    const double m0 = b.moving[u], m1 = b.moving[b.stride + u], m2 = b.moving[2 * b.stride + u];
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
        e[0] = ul - b.fixed[u];
        e[1] = vl - b.fixed[b.stride + u];
        e[2] = ur - b.fixed[2 * b.stride + u];
        e[3] = vr - b.fixed[3 * b.stride + u];
        const double om = b.omega[u];
#pragma unroll
        for (int d = 0; d < D; ++d) w[d] = om;
      } else {
        e[0] = ul - b.fixed[u];
        e[1] = vl - b.fixed[b.stride + u];
        e[2] = pc[2] - b.fixed[2 * b.stride + u];
        w[0] = w[1] = b.omega[u];
        w[2] = b.omega[b.stride + u];
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
    b.errors[u] = err;
    b.inliers[u] = inl;
    if (!use) continue;
    acc[27] += err;                                                                // :140 / :138

    // K * [wt*I3 | -2*skew(p)]  (:143-152 / :145-161); zero terms of the dense product are dropped (exact)
    const double wt = b.wt[u];
    double kj[3][6];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      kj[i][0] = K[3 * i] * wt;
      kj[i][1] = K[3 * i + 1] * wt;
      kj[i][2] = K[3 * i + 2] * wt;
      kj[i][3] = K[3 * i + 1] * (-2 * pc[2]) + K[3 * i + 2] * (-2 * -pc[1]);
      kj[i][4] = K[3 * i] * (-2 * -pc[2]) + K[3 * i + 2] * (-2 * pc[0]);
      kj[i][5] = K[3 * i] * (-2 * pc[1]) + K[3 * i + 1] * (-2 * -pc[0]);
    }
    if (KIND == 0) {
      const double il = 1 / abc[2], ir = 1 / abr[2];                               // :155-158
      const double il2 = il * il, ir2 = ir * ir;
#pragma unroll
      for (int j = 0; j < 6; ++j) {                                                // :161-177
        J[0][j] = il * kj[0][j] + (-abc[0] * il2) * kj[2][j];
        J[1][j] = il * kj[1][j] + (-abc[1] * il2) * kj[2][j];
        J[2][j] = ir * kj[0][j] + (-abr[0] * ir2) * kj[2][j];
        J[3][j] = ir * kj[1][j] + (-abr[1] * ir2) * kj[2][j];
      }
    } else {
      const double iz = 1 / pc[2], iz2 = iz * iz;                                  // :141-142
#pragma unroll
      for (int j = 0; j < 6; ++j) {                                                // :155-161
        J[0][j] = iz * kj[0][j] + (-abc[0] * iz2) * kj[2][j];
        J[1][j] = iz * kj[1][j] + (-abc[1] * iz2) * kj[2][j];
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
        double a = jw[0] * J[0][j];
#pragma unroll
        for (int d = 1; d < D; ++d) a = a + jw[d] * J[d][j];
        acc[tri(i, j)] += a;
      }
      double a = jw[0] * e[0];
#pragma unroll
      for (int d = 1; d < D; ++d) a = a + jw[d] * e[d];
      acc[21 + i] += a;
    }
  }

  // ---- warp shuffle tree, then per-block partials
  __shared__ double s_part[kThreads / 32][kAcc];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < kAcc; ++i) {
    double v = acc[i];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_part[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < kAcc) {
    double v = 0;
    for (int w = 0; w < kThreads / 32; ++w) v += s_part[w][threadIdx.x];
    b.partials[(size_t)blockIdx.x * 32 + threadIdx.x] = v;
  }
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

}  // namespace

int aligner_grid(int n, int sm_count) {
  const int blocks = (n + kThreads - 1) / kThreads;
  const int cap = sm_count * 4;   // <= 4 resident CTAs per SM at this register budget; grid-stride beyond
  return blocks < 1 ? 1 : (blocks < cap ? blocks : cap);
}

void launch_linearize(int kind, int n, const AlignerBuffers& b, const AlignerCamera& cam, const double T[12],
                      int ignore_outliers, double kernel, int grid, cudaStream_t stream) {
  Pose pose;
  for (int i = 0; i < 12; ++i) pose.T[i] = T[i];
  if (kind == 0) linearize_kernel<0><<<grid, kThreads, 0, stream>>>(n, b, cam, pose, ignore_outliers, kernel);
  else linearize_kernel<1><<<grid, kThreads, 0, stream>>>(n, b, cam, pose, ignore_outliers, kernel);
}

}  // namespace vslam
