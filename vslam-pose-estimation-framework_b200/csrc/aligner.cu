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

#include <cstdio>
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

// Warp stage of every reduction in this file: the xor-butterfly sum tree of each of the kAcc accumulators, evaluated
// TRANSPOSED -- at the step with lane distance o a lane keeps the half of its values whose index bit o matches its own
// lane bit and hands the other half to its partner, so a step moves o values instead of all of them: 31 shuffles of a
// double instead of 29 x 5.  Every accumulator still sees the sums (l, l ^ 16), then (.., l ^ 8), ... of the plain
// butterfly, and IEEE addition is commutative: the totals are bit-identical to `v += __shfl_xor_sync(v, o)`.
// Lane i < kAcc returns the warp's total of accumulator i.
__device__ __forceinline__ double warp_reduce(const double (&acc)[kAcc]) {
  const int lane = threadIdx.x & 31;
  double v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = i < kAcc ? acc[i] : 0.0;
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int j = 0; j < o; ++j) {
      const double send = upper ? v[j] : v[j + o];
      const double keep = upper ? v[j + o] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

// warp stage + cross-warp sum; the block's kAcc totals end up in threads 0..kAcc-1 (return value)
__device__ __forceinline__ double block_reduce(const double (&acc)[kAcc], double (*s_part)[kAcc]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double mine = warp_reduce(acc);
  if (lane < kAcc) s_part[warp][lane] = mine;
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

#ifdef VSLAM_GN_TIMING   // development aid: clocks of the phases of gn_step, lane 0 of warp 0 (make EXTRA_aligner="-fmad=false -DVSLAM_GN_TIMING")
__device__ __forceinline__ long long* gn_clk() {
  __shared__ long long s[16];
  return s;
}
#define GN_CLK_START() do { if ((threadIdx.x & 31) == 0) gn_clk()[15] = clock64(); } while (0)
#define GN_CLK(i) do { if ((threadIdx.x & 31) == 0) { long long* c_ = gn_clk(); const long long t_ = clock64(); c_[i] += t_ - c_[15]; c_[15] = t_; } } while (0)
// the values pass through an empty asm: what is computed from them cannot start before it, what produced them is done
#define GN_PIN6(v) asm volatile("" : "+d"(v[0]), "+d"(v[1]), "+d"(v[2]), "+d"(v[3]), "+d"(v[4]), "+d"(v[5]))
#define GN_PIN1(v) asm volatile("" : "+d"(v))
#else
#define GN_CLK_START()
#define GN_CLK(i)
#define GN_PIN6(v)
#define GN_PIN1(v)
#endif

// solve6 of gn_math.h (Eigen::FullPivLU<Matrix6>::solve) by ONE WARP on the 6 x 7 augmented matrix in SHARED memory,
// bit-identical to the single-thread form.  That form indexes its arrays with the run-time pivot position, i.e. lives in
// local memory (~8 us per Gauss-Newton round).  Two register forms were measured with clock64 at 7800 cycles per round
// each: columns spread over the lanes (~400 dependent shuffles), and everything in one thread's registers with
// compile-time indices (no memory, but ~1100 selects for the run-time row / column swaps: a single warp issues them one
// after the other).  Here a step of the elimination is a handful of instructions per lane:
//   * pivot search: lane t holds |A| of candidate t of the trailing block in row-major order (two for lanes 0..3 of the
//     first step); the maximum is three warp reductions on the bit pattern (high word, low word among the lanes that
//     hold the maximal high word, lowest candidate index among the lanes that hold the maximum: the FIRST maximum in
//     scan order, as the scalar strict `>` keeps it);
//   * row / column swap: lane c swaps its column entry of the two rows, lane r its row entry of the two columns --
//     run-time indices are free in shared memory;
//   * elimination: lane (r, c) updates A[k+1+r][k+1+c] (30 lanes in the first step), every lane dividing for itself.
// A NaN is never the maximum of the scalar scan (`fabs(a) > biggest` is false): its key is 0, below every number's.
// Every lane of the warp must call; every lane receives x[6].  A: [6][8], column 6 = right-hand side, written by the
// caller and made visible (__syncwarp) before the call.
__device__ __forceinline__ void solve6_warp(double (*A)[8], double x[6]) {
  const int lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  unsigned perm = 0x543210u;   // nibble i = perm[i]
  int rank = 6;
  double maxpivot = 0;
  // (Written without divergent branches -- clamped indices, selects, predicated stores: on this machine a divergent
  // region costs ~100 cycles of a single warp's time, measured with clock64.)
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    if (k < rank) {              // (uniform: a zero pivot ended the elimination, the break of the scalar form)
      constexpr int kSix = 6;
      const int M = kSix - k;    // the trailing block is M x M (compile-time: the loop is unrolled)
      int t = lane;
      unsigned hk, lk;
      {
        const int tc = min(lane, M * M - 1);
        const double v = fabs(A[k + tc / M][k + tc % M]);
        const bool valid = lane < M * M && v == v;
        hk = valid ? (unsigned)__double2hiint(v) + 1u : 0u;
        lk = valid ? (unsigned)__double2loint(v) : 0u;
      }
      if (M * M > 32) {          // candidates 32..35 of the first step: later in scan order, they win when larger
        const int t2 = min(lane + 32, M * M - 1);
        const double v = fabs(A[k + t2 / M][k + t2 % M]);
        const bool valid = lane + 32 < M * M && v == v;
        const unsigned h2 = valid ? (unsigned)__double2hiint(v) + 1u : 0u, l2 = valid ? (unsigned)__double2loint(v) : 0u;
        const bool take = h2 > hk || (h2 == hk && l2 > lk);
        hk = take ? h2 : hk;
        lk = take ? l2 : lk;
        t = take ? t2 : t;
      }
      GN_CLK(0);
      const unsigned mh = __reduce_max_sync(full, hk);
      const unsigned ml = __reduce_max_sync(full, hk == mh ? lk : 0u);
      const int tw = (int)__reduce_min_sync(full, (hk == mh && lk == ml) ? (unsigned)t : 64u);
      // (mh == 0: every candidate is a NaN: the scalar scan keeps -1 and (k, k))
      const double biggest = mh != 0 ? __hiloint2double((int)(mh - 1u), (int)ml) : -1.0;
      const int pr = mh != 0 ? k + tw / M : k;
      const int pc = mh != 0 ? k + tw % M : k;
      GN_CLK(1);
      if (biggest == 0) {        // (uniform)
        rank = k;
      } else {
        maxpivot = biggest > maxpivot ? biggest : maxpivot;
        {                        // rows k and pr (the right-hand side in column 6 follows); pr == k rewrites the row
          const int c = min(lane, 6);
          const double a = A[k][c], b = A[pr][c];
          __syncwarp();
          A[k][c] = b;
          A[pr][c] = a;
        }
        __syncwarp();
        {                        // columns k and pc (every row: the finished rows of U follow the permutation)
          const int r = min(lane, 5);
          const double a = A[r][k], b = A[r][pc];
          __syncwarp();
          A[r][k] = b;
          A[r][pc] = a;
          const unsigned pk = (perm >> (4 * k)) & 15u, pp = (perm >> (4 * pc)) & 15u;
          perm = (perm & ~((15u << (4 * k)) | (15u << (4 * pc)))) | (pp << (4 * k)) | (pk << (4 * pc));
        }
        __syncwarp();
        GN_CLK(2);
        {
          const int r = lane / 6, c = lane - 6 * r;
          const bool mine = k + 1 + r < 6 && k + 1 + c < 7;
          const int i = min(k + 1 + r, 5), j = min(k + 1 + c, 6);
          const double f = A[i][k] / A[k][k];
          const double v = A[i][j] - f * A[k][j];
          __syncwarp();
          if (mine) A[i][j] = v;
        }
        __syncwarp();
        GN_CLK(3);
      }
    }
  }
  // Eigen::FullPivLU::rank(): only pivots above |largest pivot| * epsilon * size are used by solve()
  GN_PIN1(maxpivot);
  GN_CLK(7);
  {
    int r = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) r += (i < rank && fabs(A[i][i]) > maxpivot * (2.220446049250313e-16 * 6)) ? 1 : 0;
    rank = r;
  }
  // back substitution, every lane for itself (broadcast reads); full rank -- the regular case -- as straight-line code
  double y[6];
  if (rank == 6) {
#pragma unroll
    for (int i = 5; i >= 0; --i) {
      double sum = A[i][6];
#pragma unroll
      for (int j = i + 1; j < 6; ++j) sum = sum - A[i][j] * y[j];
      y[i] = sum / A[i][i];
    }
  } else {
#pragma unroll
    for (int i = 0; i < 6; ++i) y[i] = 0;
#pragma unroll
    for (int i = 5; i >= 0; --i) {
      if (i < rank) {
        double sum = A[i][6];
#pragma unroll
        for (int j = 0; j < 6; ++j)
          if (j > i && j < rank) sum = sum - A[i][j] * y[j];
        y[i] = sum / A[i][i];
      }
    }
  }
  // x[perm[i]] = y[i]
#pragma unroll
  for (int t = 0; t < 6; ++t) x[t] = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int t = 0; t < 6; ++t) x[t] = (int)((perm >> (4 * i)) & 15u) == t ? y[i] : x[t];
  GN_PIN6(x);
  GN_CLK(4);
}

// oneRound (:190-207 / :174-191) + the converge state machine (:213-247 / :197-233) for one finished linearisation:
// damped system, full-pivot solve, pose update, bookkeeping in *ctl (global memory in the grid kernel, shared memory in
// the cluster kernel).  Called by the WHOLE first warp of the block (the solve is warp-cooperative).
// `s_A` is the solver's [6][8] scratch in shared memory.
template <bool kControlInGlobalMemory>
__device__ __forceinline__ void gn_step(GnControl* ctl, const double* s_sys, const double* s_T, double (*s_A)[8], int n,
                                        const GnParams& p) {
  const int lane = threadIdx.x & 31;
  GN_CLK_START();
  {   // H as the single-thread form builds it: both triangles from the packed upper one, then the damping on the
      // diagonal; the solver's working copy carries -b as a seventh column
    for (int idx = lane; idx < 36; idx += 32) {
      const int i = idx / 6, j = idx - 6 * i;
      double v = s_sys[tri(min(i, j), max(i, j))];
      if (i == j) v += p.damping * n;
      ctl->H[idx] = v;
      s_A[i][j] = v;
    }
    if (lane < 6) s_A[lane][6] = -s_sys[21 + lane];
    __syncwarp();
  }
  double dx[6];
  solve6_warp(s_A, dx);
  double T[12];   // (every lane: the update is a few dozen operations, the pose is written by 12 lanes at once)
#pragma unroll
  for (int i = 0; i < 12; ++i) T[i] = s_T[i];
  GN_PIN6(T);
  GN_CLK(8);
  apply_update(dx, T);
  GN_PIN6(T);
  {
    double* T6 = T + 6;
    GN_PIN6(T6);
  }
  GN_CLK(9);
  {   // lane i < 12 stores T[i] (a select chain: twelve divergent one-lane stores cost 2000 cycles)
    double mine = T[0];
#pragma unroll
    for (int i = 1; i < 12; ++i) mine = lane == i ? T[i] : mine;
    if (lane < 12) ctl->T[lane] = mine;
  }
  GN_CLK(5);
  if (lane != 0) return;
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
  if (kControlInGlobalMemory) __threadfence();   // (other blocks read it after the grid barrier)
  ctl->done = done;
  GN_CLK(6);
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
  __shared__ double s_A[6][8];          // working copy of the damped system, column 6 = -b
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
      if (threadIdx.x < 32) gn_step<true>(ctl, s_sys, s_T, s_A, n, p);
    }
    grid.sync();
    if (__ldcg(&ctl->done)) break;
  }
}

// The same loop for the problem sizes of a tracked frame (one correspondence per thread): ONE thread-block cluster
// instead of a cooperative grid.  The CTAs of the cluster are the blocks of the grid version -- same per-point code, same
// per-block partials, same order of the final sum, so pose and round count stay bit-identical -- but
//  * every thread keeps its ONE correspondence in registers across the rounds;
//  * a round has ONE barrier, a hardware cluster barrier: every block publishes its partial sums in its own shared
//    memory (double-buffered by round parity, so a block that runs ahead cannot overwrite what a slower block still
//    reads), then EVERY block gathers all partials through distributed shared memory in block order and runs the
//    damped solve, the pose update and the convergence state machine itself on its own copy of the control block --
//    redundant, deterministic and identical in every block, instead of block 0 solving while the others wait at a
//    second barrier and then fetch the pose from it;
//  * blocks without a correspondence exit at once (their partials would be +0.0: skipping them leaves every sum
//    unchanged; a cluster barrier waits for the non-exited threads only).
// Launched as a plain kernel with a cluster dimension.
template <int KIND>
__global__ void __launch_bounds__(kThreads, 1) converge_cluster_kernel(int n, AlignerBuffers b, AlignerCamera cam, GnParams p,
                                                                    GnControl* __restrict__ ctl,
                                                                    const int32_t* __restrict__ n_device, FrameFill fill,
                                                                    FramePrune prune) {
  constexpr int D = KIND == 0 ? 4 : 3;
  constexpr int W = KIND == 0 ? 1 : 2;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  // fused frame (captured graph): the correspondence count is what track() left in device memory; nothing to align
  // without tracks (pose_tracker_3d.cpp:355: the tracker only optimises a frame that has points)
  if (n_device) {
    n = *n_device;
    const bool run = n > 0 && n <= (int)(gridDim.x * kThreads);
    if (fill.tracks && rank == 0 && !run) {   // the frame still publishes a control block: the prior, no rounds
      if (threadIdx.x < 12) ctl->T[threadIdx.x] = fill.state->T_prior[threadIdx.x];
      if (threadIdx.x < 36) ctl->H[threadIdx.x] = 0.0;
      if (threadIdx.x < kAcc) b.system[threadIdx.x] = 0.0;
      if (threadIdx.x == 0) {
        ctl->total_error_previous = 0.0;
        ctl->rounds = ctl->phase = ctl->iteration = ctl->ignore = ctl->converged = ctl->done = 0;
        fill.state->overflow = n > (int)(gridDim.x * kThreads) ? 1 : 0;
        fill.state->n_kept = 0;
        fill.state->inliers_only = 0;
      }
    }
    if (!run) return;
  }
  const int n_active = (n + kThreads - 1) / kThreads;     // blocks that hold a correspondence
  if (rank >= n_active) return;
  __shared__ double s_part[kThreads / 32][kAcc];
  __shared__ double s_total[2][32];     // this block's partial sums of the even / odd rounds, read by every block
  __shared__ double s_T[12];
  __shared__ double s_sys[32];
  __shared__ double s_A[6][8];          // working copy of the damped system, column 6 = -b
  __shared__ GnControl s_ctl;           // every block advances its own copy (identically)

  if (fill.tracks) {                   // converge() of a fused frame starts from the motion prior (stereouv_aligner.cpp:213-216)
    if (threadIdx.x < 12) s_ctl.T[threadIdx.x] = fill.state->T_prior[threadIdx.x];
    if (threadIdx.x == 0) {
      s_ctl.total_error_previous = 0.0;
      s_ctl.rounds = s_ctl.phase = s_ctl.iteration = s_ctl.ignore = s_ctl.converged = s_ctl.done = 0;
      if (rank == 0) {
        fill.state->overflow = 0;
        fill.state->n_kept = 0;
        fill.state->inliers_only = 0;
      }
    }
  } else {
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
  if (KIND == 0 && fill.tracks) {
    if (mine) {                        // StereoUVAligner::initialize for correspondence u (stereouv_aligner.cpp:26-64)
      const TrackRecord* t = fill.tracks + u;
      const PreviousPoint* pp = fill.previous + t->index_previous;
      m[0] = pp->camera[0];            // :52-55 previous->cameraCoordinatesLeft()
      m[1] = pp->camera[1];
      m[2] = pp->camera[2];
      fx[0] = (double)t->xl;           // :36-39
      fx[1] = (double)t->yl;
      fx[2] = (double)t->xr;
      fx[D - 1] = (double)t->yr;
      om[0] = 1.0;                     // :28 setIdentity
      {                                // :43-51 a landmark estimate is preferred, its information grows with the updates
        const LandmarkEstimate le = fill.estimates[t->index_previous];
        if (le.information_scale != 0.0) {
          m[0] = le.camera[0];
          m[1] = le.camera[1];
          m[2] = le.camera[2];
          om[0] = le.information_scale;   // identity * (1 + log(numberOfUpdates)), evaluated by the host
        }
      }
      wt = 1.0;
      if (fill.inverse_depth_weight) { // :59-63 std::min(max_depth / depth, 1.0)
        const double ratio = fill.max_reliable_depth / t->camera[2];
        wt = 1.0 < ratio ? 1.0 : ratio;
      }
      fill.track_length[u] = pp->reserved;   // trackLength() of the previous point (frame_point.cpp:27)
    }
  } else if (mine) {
#pragma unroll
    for (int d = 0; d < 3; ++d) m[d] = b.moving[d * b.stride + u];
#pragma unroll
    for (int d = 0; d < D; ++d) fx[d] = b.fixed[d * b.stride + u];
#pragma unroll
    for (int d = 0; d < W; ++d) om[d] = b.omega[d * b.stride + u];
    wt = b.wt[u];
  }
  __syncthreads();

  double err = -1.0;
  uint8_t inl = 0;
#ifdef VSLAM_GN_TIMING   // phase clocks of block 0 (development aid: make EXTRA_aligner="-fmad=false -DVSLAM_GN_TIMING")
  if (threadIdx.x < 16) gn_clk()[threadIdx.x] = 0;
  long long t_phase[6] = {0, 0, 0, 0, 0, 0}, t_mark = clock64();
#define GN_MARK(i) { const long long t_now = clock64(); t_phase[i] += t_now - t_mark; t_mark = t_now; }
#else
#define GN_MARK(i)
#endif
  for (int parity = 0;; parity ^= 1) {
    if (threadIdx.x < 12) s_T[threadIdx.x] = s_ctl.T[threadIdx.x];
    const int ignore_outliers = s_ctl.ignore;
    __syncthreads();
    GN_MARK(0)
    double acc[kAcc];
#pragma unroll
    for (int i = 0; i < kAcc; ++i) acc[i] = 0.0;
    if (mine) accumulate_point<KIND>(m[0], m[1], m[2], fx, om, wt, s_T, cam, ignore_outliers, p.kernel, acc, err, inl);
    GN_MARK(1)
    const double total = block_reduce(acc, s_part);
    if (threadIdx.x < kAcc) s_total[parity][threadIdx.x] = total;
    GN_MARK(2)
    cluster.sync();                     // every block's partial of this round is in its shared memory
    GN_MARK(3)

    if (n_active <= kThreads / 32) {
      // At most one block per slice: the two-level sum of linearize_kernel (slices of every 8th block, then the 8 slices)
      // degenerates to the chain 0 + t_0 + t_1 + ... over the blocks -- the empty slices add +0.0, which changes nothing --
      // so the first kAcc threads (all in warp 0, which goes on to solve) add the blocks' partials directly: no block
      // barrier, no trip through s_part.  Everybody else meets them at the barrier behind gn_step.
      if (threadIdx.x < kAcc) {
        double t[kThreads / 32];
#pragma unroll
        for (int blk = 0; blk < kThreads / 32; ++blk)
          t[blk] = blk < n_active ? cluster.map_shared_rank(&s_total[parity][0], blk)[threadIdx.x] : 0.0;
        double s = 0;
#pragma unroll
        for (int blk = 0; blk < kThreads / 32; ++blk) s += t[blk];
        s_sys[threadIdx.x] = s;
      }
      if (threadIdx.x < 32) __syncwarp();
    } else {  // the block partials in block order, 8 interleaved slices per value -- as in linearize_kernel
      const int j = threadIdx.x & 31, part = threadIdx.x >> 5;
      double v = 0;
      if (j < kAcc)
        for (int blk = part; blk < n_active; blk += kThreads / 32) v += cluster.map_shared_rank(&s_total[parity][0], blk)[j];
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
    GN_MARK(4)
    if (threadIdx.x < 32) gn_step<false>(&s_ctl, s_sys, s_T, s_A, n, p);
    GN_MARK(5)
    __syncthreads();
    if (s_ctl.done) break;
  }
#ifdef VSLAM_GN_TIMING
  if (rank == 0 && threadIdx.x == 0)
    printf("gn_step clocks: build + candidates %lld | pivot reductions %lld | swaps %lld | elimination %lld | (pin) %lld | rank + back "
           "substitution %lld | load T %lld | apply_update %lld | store T %lld | state machine %lld\n", gn_clk()[0], gn_clk()[1],
           gn_clk()[2], gn_clk()[3], gn_clk()[7], gn_clk()[4], gn_clk()[8], gn_clk()[9], gn_clk()[5], gn_clk()[6]);
  if (rank == 0 && threadIdx.x == 0)
    printf("gn timing (cycles, %d rounds): load T %lld | accumulate %lld | block reduce %lld | cluster barrier %lld | gather %lld | "
           "solve + update %lld\n", s_ctl.rounds, t_phase[0], t_phase[1], t_phase[2], t_phase[3], t_phase[4], t_phase[5]);
#endif
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
  if (!prune.tracked) {
    cluster.sync();                     // no block may exit while another still reads its partials of the last round
    return;
  }
  // ---- _prunePoints (pose_tracker_3d.cpp:437-472) over the cluster: keep = inlier when the average error
  // (total error / correspondences, base_aligner.h:46) is below the kernel, else error != -1 && error < 100 kernel; the
  // kept records keep their order: block scan, block totals through distributed shared memory, one cluster barrier
  // (every record is read before it, every write comes after it: the compaction is in place)
  __shared__ int s_keep_warp[kThreads / 32];
  __shared__ int s_keep_total;
  const bool inliers_only = s_sys[27] / n < prune.error_kernel;     // (n > 0 here)
  const bool keep = mine && (inliers_only ? inl != 0 : (err != -1.0 && err < 100 * prune.error_kernel));
  uint4 rec0 = make_uint4(0, 0, 0, 0), rec1 = rec0;
  if (keep) {
    const uint4* src = reinterpret_cast<const uint4*>(prune.tracked + u);
    rec0 = src[0];
    rec1 = src[1];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) s_keep_warp[warp] = __popc(bal);
  __syncthreads();
  int before = 0, block_total = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) {
    const int c = s_keep_warp[w];
    if (w < warp) before += c;
    block_total += c;
  }
  if (threadIdx.x == 0) s_keep_total = block_total;
  cluster.sync();                       // totals published, records read, partials of the last round no longer needed
  int all = 0;
  for (int r = 0; r < n_active; ++r) {
    const int c = *cluster.map_shared_rank(&s_keep_total, r);
    if (r < rank) before += c;
    all += c;
  }
  if (mine) {
    const int pos = before + __popc(bal & ((1u << lane) - 1u));
    prune.kept_pos[u] = keep ? pos : -1;
    if (keep) {
      uint4* dst = reinterpret_cast<uint4*>(prune.tracked + pos);
      dst[0] = rec0;
      dst[1] = rec1;
    }
  }
  if (rank == 0 && threadIdx.x == 0) {
    prune.state->n_kept = all;
    prune.state->inliers_only = inliers_only;
  }
  cluster.sync();                       // no block may exit while another still reads its total
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
  const double mine = warp_reduce(acc);   // lane i ends up with the warp's total i
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
    const FrameFill no_fill = {nullptr, nullptr, nullptr, nullptr, nullptr, 0.0, 0};
    const FramePrune no_prune = {nullptr, nullptr, nullptr, 0.0};
    const cudaError_t e =
        kind == 0 ? cudaLaunchKernelEx(&cfg, converge_cluster_kernel<0>, n_arg, b_arg, cam_arg, p_arg, ctl, no_device_count, no_fill, no_prune)
                  : cudaLaunchKernelEx(&cfg, converge_cluster_kernel<1>, n_arg, b_arg, cam_arg, p_arg, ctl, no_device_count, no_fill, no_prune);
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
                                  const int32_t* n_device, int cluster_blocks, const FrameFill& fill, const FramePrune& prune,
                                  cudaStream_t stream) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr;
  cluster_config(&cfg, &attr, cluster_blocks, stream);
  return cudaLaunchKernelEx(&cfg, converge_cluster_kernel<0>, 0, b, cam, p, ctl, n_device, fill, prune);
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
