// track.cu -- StereoFramePointGenerator::track and ::recoverPoints on the device
// (reference src/framepoint_generation/stereo_framepoint_generator.cpp:464-681, 683-869 and
// IntensityFeatureMatcher::getMatchingFeatureInRectangularRegion, intensity_feature_matcher.cpp:81-148).
//
// track() is sequential over the previous frame's points: a point that is tracked AND triangulated removes its left
// feature, its right feature and the right features in its parallax range from the lattices, so later points cannot
// pick them.  Removing a feature changes another point's result only if it was that point's chosen left or right
// feature (an argmin with first-wins ties does not move when a loser disappears).  Hence:
//
//   K9  track_search_kernel   one warp per previous point, all points in parallel against the UNCONSUMED lattices:
//                             projection, rectangular window search (popc-256, warp argmin), corrected right window.
//   K10 track_resolve_kernel  one CTA per frame; ordered fixed-point iteration.  Every tentative success claims the
//                             features it removes with atomicMin(point index): feature c is gone FOR POINT i iff
//                             claim[c] < i, which is exactly the lattice the sequential loop shows to point i once the
//                             results of the points below i are right.  A point is DIRTY when a lower point claims one
//                             of the (at most two) features its result depends on.  Each round rebuilds the claims from
//                             the current results and recomputes, one warp per point, every point that has ever been
//                             dirty against claim[.] < i; the rounds stop when nothing changes.  By induction over the
//                             point index the fixed point is unique and equals the sequential result (point 0 never
//                             depends on anyone; point i is right as soon as 0..i-1 are), and it is reached after as
//                             many rounds as the longest dependency chain -- a handful, because the dominant conflict
//                             (a track removes the right features in its parallax range, :611-620) stays within one
//                             image row.  The kernel then emits tracks / lost points in order, the pruned flags and
//                             the bin pre-load records.
//
// The lattice of the reference is replaced by the (row, col)-sorted feature arrays + CSR row pointers: a window's rows
// are one contiguous index range, scanned in the reference's row-major order (ties resolve to the lowest index).
#include <algorithm>
#include <cstdio>

#include "kernels.cuh"
#include "stereo_device.cuh"

namespace vslam {

namespace {

constexpr int kStatusSkipped = 0;   // left the loop body through `continue`: neither tracked nor lost
constexpr int kStatusLost = 1;      // reached :660-663 without a track
constexpr int kStatusTracked = 2;

struct TrackView {
  const int32_t* rpl;
  const int32_t* rpr;
  const uint32_t* xyl;
  const uint32_t* xyr;
  const uint4* dl;
  const uint4* dr;
  const uint8_t* gone_l;   // features removed before track() (pruned flags at entry)
  const uint8_t* gone_r;
  const int32_t* claim_l;  // when non-null: feature f is removed for point `self` iff claim[f] < self
  const int32_t* claim_r;
  int self;
  int rows, cols;
  double fx, fy, cx, cy, bx;
  double threshold_triangulation;   // _current_maximum_descriptor_distance_triangulation
  double min_disparity;
};

struct Projection {
  int32_t col_l, row_l, col_r, row_r;
  float right_x, right_y;
};

// The claims and the tentative results of track_resolve_kernel live in shared memory when they fit and in global memory
// otherwise; both are written by other threads between two barriers (atomics at the L2 / plain stores), so every read
// goes to the memory itself: volatile loads through the generic address (global: past the L1; shared: as they are).
__device__ __forceinline__ int32_t load_claim(const int32_t* p) { return *reinterpret_cast<const volatile int32_t*>(p); }
__device__ __forceinline__ int4 load_result(const int4* p) {
  int4 r;
  asm volatile("ld.volatile.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// intensity_feature_matcher.cpp:81-148, warp-cooperative.  Returns the sorted index of the chosen feature or -1;
// *distance = descriptor_distance_best_ of the chosen feature.
__device__ int search_region(const int32_t* __restrict__ row_ptr, const uint32_t* __restrict__ xy,
                             const uint4* __restrict__ desc, const uint8_t* gone, const int32_t* claim, int self,
                             int row_reference, int col_reference,
                             const uint4& q0, const uint4& q1, int row_start, int row_end, int col_start, int col_end,
                             double maximum_distance, bool by_appearance, int* distance) {
  const int lane = threadIdx.x & 31;
  if (row_start >= row_end || col_start >= col_end) return -1;
  const int f0 = row_ptr[row_start], f1 = row_ptr[row_end];
  unsigned best = 0xffffffffu;
  for (int f = f0 + lane; f < f1; f += 32) {
    // (the position and the removed flag are fetched together: one trip to memory instead of two in a row)
    const uint32_t q = xy[f];
    const bool removed = claim ? load_claim(claim + f) < self : gone[f] != 0;   // feature_lattice[row][col] == nullptr
    const int col = (int)(q & 0xffffu);
    if (col < col_start || col >= col_end || removed) continue;
    const int d = popc256(q0, q1, desc[2 * f], desc[2 * f + 1]);
    if (!((double)d < maximum_distance)) continue;                   // :103 / :121 (strict, against the double limit)
    unsigned key;
    if (by_appearance) {
      key = ((unsigned)d << 16) | (unsigned)f;                       // :103-107 first strict minimum in scan order
    } else {
      const int dr = row_reference - (int)(q >> 16), dc = col_reference - col;
      const unsigned pd = (unsigned)(dr * dr + dc * dc);             // :124-126
      if (pd >= 10000u) continue;                                    // :115, :129
      key = (pd << 16) | (unsigned)f;
    }
    best = min(best, key);
  }
  best = __reduce_min_sync(0xffffffffu, best);
  if (best == 0xffffffffu) return -1;
  const int f = (int)(best & 0xffffu);
  *distance = by_appearance ? (int)(best >> 16) : popc256(q0, q1, desc[2 * f], desc[2 * f + 1]);
  return f;
}

// one iteration of the loop :494-664 for previous point `pp` against the lattices as `v` shows them.
// result: {left feature | -1, right feature | -1, descriptor_distance_best, status}
__device__ int4 track_point(const TrackView& v, const TrackParams& tp, const PreviousPoint* __restrict__ pp,
                            Projection* proj) {
  const double X = pp->camera[0], Y = pp->camera[1], Z = pp->camera[2];
  double pc[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)                                                                    // :496-498
    pc[i] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(tp.T[4 * i], X), __dmul_rn(tp.T[4 * i + 1], Y)),
                                __dmul_rn(tp.T[4 * i + 2], Z)), tp.T[4 * i + 3]);
  const double il0 = __dadd_rn(__dmul_rn(v.fx, pc[0]), __dmul_rn(v.cx, pc[2]));                  // :501-502
  const double il1 = __dadd_rn(__dmul_rn(v.fy, pc[1]), __dmul_rn(v.cy, pc[2]));
  const double il2 = pc[2];
  const int32_t col_l = to_i32(__ddiv_rn(il0, il2)), row_l = to_i32(__ddiv_rn(il1, il2));        // :503-506
  proj->col_l = col_l;
  proj->row_l = row_l;
  if (col_l < 0 || col_l > v.cols || row_l < 0 || row_l > v.rows) return make_int4(-1, -1, 0, kStatusSkipped);

  const int D = tp.distance_pixels;
  const uint4* ql = reinterpret_cast<const uint4*>(pp->descriptor_left);
  const uint4 q0 = ql[0], q1 = ql[1];
  int distance = 0;
  const int fl = search_region(v.rpl, v.xyl, v.dl, v.gone_l, v.claim_l, v.self, row_l, col_l, q0, q1, max(row_l - D, 0),   // :520-538
                               min(row_l + D + 1, v.rows), max(col_l - D, 0), min(col_l + D + 1, v.cols),
                               tp.max_distance_tracking, tp.by_appearance != 0, &distance);
  if (fl < 0) return make_int4(-1, -1, 0, kStatusLost);

  const uint32_t pl = v.xyl[fl];
  const int col_fl = (int)(pl & 0xffffu), row_fl = (int)(pl >> 16);
  const float error_x = __fsub_rn((float)col_l, (float)col_fl);                                  // :543-545
  const float error_y = __fsub_rn((float)row_l, (float)row_fl);
  const double ir0 = __dadd_rn(il0, v.bx), ir1 = __dadd_rn(il1, 0.0), ir2 = __dadd_rn(il2, 0.0); // :549
  const double rx = __ddiv_rn(ir0, ir2), ry = __ddiv_rn(ir1, ir2);
  const int32_t col_r = to_i32(__dsub_rn(rx, (double)error_x));                                  // :550-555
  const int32_t row_r = to_i32(__dsub_rn(ry, (double)error_y));
  proj->col_r = col_r;
  proj->row_r = row_r;
  proj->right_x = (float)rx;
  proj->right_y = (float)ry;
  if (col_r < 0 || col_r > v.cols || row_r < 0 || row_r > v.rows) return make_int4(fl, -1, 0, kStatusSkipped);

  const int e = (int)fabs((double)pp->epipolar_offset);                                          // :568-569
  const uint4 l0 = v.dl[2 * fl], l1 = v.dl[2 * fl + 1];
  const int fr = search_region(v.rpr, v.xyr, v.dr, v.gone_r, v.claim_r, v.self, row_r, col_r, l0, l1, max(row_r - e, 0),   // :570-590
                               min(row_r + e + 1, v.rows), max(col_r - D, 0), min(col_r + D + 1, col_fl),
                               v.threshold_triangulation, true, &distance);
  if (fr < 0) return make_int4(fl, -1, 0, kStatusLost);
  const int col_fr = (int)(v.xyr[fr] & 0xffffu);
  if ((double)(col_fl - col_fr) < v.min_disparity) return make_int4(fl, fr, distance, kStatusSkipped);   // :597-600
  const uint4* qr = reinterpret_cast<const uint4*>(pp->descriptor_right);
  if ((double)popc256(v.dr[2 * fr], v.dr[2 * fr + 1], qr[0], qr[1]) > tp.max_distance_tracking)  // :603-607
    return make_int4(fl, fr, distance, kStatusSkipped);
  return make_int4(fl, fr, distance, kStatusTracked);
}

__device__ __forceinline__ TrackView make_view(const Geometry& g, const StereoParams& sp, const int32_t* row_ptr,
                                               const uint32_t* kp_xy, const uint8_t* desc, const int32_t* n_desc,
                                               const uint8_t* gone_l, const uint8_t* gone_r) {
  TrackView v;
  v.rpl = row_ptr;
  v.rpr = row_ptr + (g.rows + 1);
  v.xyl = kp_xy;
  v.xyr = kp_xy + g.cap;
  v.dl = reinterpret_cast<const uint4*>(desc);
  v.dr = reinterpret_cast<const uint4*>(desc + (size_t)g.cap * kDescBytes);
  v.gone_l = gone_l;
  v.gone_r = gone_r;
  v.claim_l = nullptr;
  v.claim_r = nullptr;
  v.self = 0;
  v.rows = g.rows;
  v.cols = g.cols;
  v.fx = sp.fx; v.fy = sp.fy; v.cx = sp.cx; v.cy = sp.cy; v.bx = sp.bx;
  v.threshold_triangulation = triangulation_threshold(sp, n_desc[0]);
  v.min_disparity = sp.min_disparity;
  return v;
}

// K9: every previous point against the lattices as initialize() (or the caller) left them
__global__ void __launch_bounds__(256) track_search_kernel(Geometry g, StereoParams sp, TrackParams tp,
                                                           const int32_t* row_ptr, const uint32_t* kp_xy,
                                                           const uint8_t* desc, const int32_t* n_desc,
                                                           const uint8_t* gone_l, const uint8_t* gone_r,
                                                           const PreviousPoint* __restrict__ previous, int n_previous,
                                                           int4* __restrict__ tentative,
                                                           const FrameStepState* __restrict__ step) {
  const int u = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (step) {   // fused frame: the point count and the motion prior live in device memory (n_previous = capacity)
    n_previous = min(n_previous, step->n_previous);
    if (u >= n_previous) return;
#pragma unroll
    for (int i = 0; i < 12; ++i) tp.T[i] = step->T_prior[i];
  }
  if (u >= n_previous) return;
  const TrackView v = make_view(g, sp, row_ptr, kp_xy, desc, n_desc, gone_l, gone_r);
  Projection proj;
  const int4 r = track_point(v, tp, previous + u, &proj);
  if ((threadIdx.x & 31) == 0) tentative[u] = r;
}

// right features a track removes: the match and the parallax range behind it on its row (:611-620, :646-651)
template <typename F>
__device__ __forceinline__ void for_each_consumed_right(const TrackView& v, int fl, int fr, F&& f) {
  const int col_fl = (int)(v.xyl[fl] & 0xffffu);
  const int row_fr = (int)(v.xyr[fr] >> 16);
  const int end = v.rpr[row_fr + 1];
  f(fr);
  for (int s = fr + 1; s < end && (int)(v.xyr[s] & 0xffffu) < col_fl; ++s) f(s);
}

constexpr int kResolveThreads = 1024;

// exclusive scan of a 0/1 flag over the block; total in *total
__device__ __forceinline__ int block_scan_flag(bool flag, int* s_red, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, flag);
  if (lane == 0) s_red[warp] = __popc(bal);
  __syncthreads();
  const int mine = s_red[lane];
  int inc = mine;
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  const int warp_offset = __shfl_sync(0xffffffffu, inc - mine, warp);
  *total = __shfl_sync(0xffffffffu, inc, 31);
  __syncthreads();
  return warp_offset + __popc(bal & ((1u << lane) - 1u));
}

#ifdef VSLAM_TRACK_TIMING   // development aid: phase clocks of track_resolve_kernel (make EXTRA_track="-fmad=false -DVSLAM_TRACK_TIMING")
#define TR_MARK(i) { if (threadIdx.x == 0) { const long long t_now = clock64(); t_phase[i] += t_now - t_mark; t_mark = t_now; } }
#else
#define TR_MARK(i)
#endif

// K10.  Dynamic shared memory (layout chosen by launch_track, 0 = the global scratch is used): the two claim arrays
// [2][smem_claims] and the tentative results + the worklist of the dirty points [smem_points] -- a round then runs out
// of shared memory instead of paying an L2 round trip per phase (reset, atomicMin, dirty test, recompute).
__global__ void __launch_bounds__(kResolveThreads) track_resolve_kernel(
    Geometry g, StereoParams sp, TrackParams tp, const int32_t* row_ptr, const uint32_t* kp_xy, const uint8_t* desc,
    const int32_t* n_desc, uint8_t* gone_l, uint8_t* gone_r, const PreviousPoint* __restrict__ previous,
    int n_previous, int4* tentative, int32_t* claim_l, int32_t* claim_r, int4* __restrict__ final,
    int32_t* __restrict__ lost, uint8_t* __restrict__ ever_dirty, int32_t* __restrict__ stats,
    const FrameStepState* __restrict__ step, int smem_claims, int smem_points) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ int s_red[32];
  __shared__ int s_acc[2];
  const int tid = threadIdx.x;
  if (step) {   // fused frame: see track_search_kernel
    n_previous = min(n_previous, step->n_previous);
#pragma unroll
    for (int i = 0; i < 12; ++i) tp.T[i] = step->T_prior[i];
  }
  const TrackView v = make_view(g, sp, row_ptr, kp_xy, desc, n_desc, gone_l, gone_r);
  const int n_l = n_desc[0], n_r = n_desc[1];

  __shared__ int s_count, s_changed;
  int32_t* worklist = lost;          // scratch until the ordered output below fills it
  if (smem_claims) {
    claim_l = reinterpret_cast<int32_t*>(s_dyn);
    claim_r = claim_l + smem_claims;
  }
  if (smem_points) {                 // (n_previous <= smem_points: launch_track)
    int4* s_tentative = reinterpret_cast<int4*>(s_dyn + ((sizeof(int32_t) * 2 * (size_t)smem_claims + 15) & ~(size_t)15));
    for (int u = tid; u < n_previous; u += kResolveThreads) s_tentative[u] = tentative[u];
    tentative = s_tentative;
    worklist = reinterpret_cast<int32_t*>(s_tentative + smem_points);
  }
  TrackView vc = v;   // the lattice as point `self` sees it: claims of lower points
  vc.claim_l = claim_l;
  vc.claim_r = claim_r;
  for (int u = tid; u < n_previous; u += kResolveThreads) ever_dirty[u] = 0;
  if (tid < 2) s_acc[tid] = 0;

#ifdef VSLAM_TRACK_TIMING
  long long t_phase[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t_mark = clock64();
  int n_rounds = 0, work_total = 0;
#endif
  for (int round = 0; round <= n_previous; ++round) {
#ifdef VSLAM_TRACK_TIMING
    ++n_rounds;
#endif
    // claims of the current results; a feature that was gone at entry is gone for everybody (claim -1)
    for (int i = tid; i < n_l; i += kResolveThreads) claim_l[i] = gone_l[i] ? -1 : INT32_MAX;
    for (int i = tid; i < n_r; i += kResolveThreads) claim_r[i] = gone_r[i] ? -1 : INT32_MAX;
    if (tid == 0) s_count = 0, s_changed = 0;
    __syncthreads();
    TR_MARK(0)
    for (int u = tid; u < n_previous; u += kResolveThreads) {
      const int4 t = load_result(tentative + u);
      if (t.w != kStatusTracked) continue;
      atomicMin(&claim_l[t.x], u);
      for_each_consumed_right(v, t.x, t.y, [&](int s) { atomicMin(&claim_r[s], u); });
    }
    __syncthreads();
    TR_MARK(1)
    // points whose result depends on a feature a lower point removes join the set that is recomputed every round
    for (int u = tid; u < n_previous; u += kResolveThreads) {
      const int4 t = load_result(tentative + u);
      bool d = ever_dirty[u] != 0;
      if (!d && ((t.x >= 0 && load_claim(claim_l + t.x) < u) || (t.y >= 0 && load_claim(claim_r + t.y) < u))) {
        d = true;
        ever_dirty[u] = 1;
        s_changed = 1;
      }
      if (d) worklist[atomicAdd(&s_count, 1)] = u;
    }
    __syncthreads();
    TR_MARK(2)
    const int n_work = s_count;
#ifdef VSLAM_TRACK_TIMING
    work_total += n_work;
#endif
    if (n_work == 0) break;
    for (int k = tid >> 5; k < n_work; k += kResolveThreads / 32) {
      const int u = *reinterpret_cast<volatile int32_t*>(worklist + k);
      vc.self = u;
      Projection proj;
      const int4 t = track_point(vc, tp, previous + u, &proj);
      if ((tid & 31) == 0) {
        const int4 old = load_result(tentative + u);
        if (old.x != t.x || old.y != t.y || old.z != t.z || old.w != t.w) {
          tentative[u] = t;
          s_changed = 1;
        }
      }
    }
    __syncthreads();
    TR_MARK(3)
    if (!s_changed) break;   // every point satisfies its own equation: the sequential result
    __syncthreads();
  }
  TR_MARK(4)
  // the claims of the final results == matched_indices_left / _right of :646-651 (+ the parallax ranges :611-620)
  for (int i = tid; i < n_l; i += kResolveThreads) gone_l[i] = load_claim(claim_l + i) != INT32_MAX;
  for (int i = tid; i < n_r; i += kResolveThreads) gone_r[i] = load_claim(claim_r + i) != INT32_MAX;
  __syncthreads();

  // ordered positions: tracks (:623-643) and lost points (:660-663) keep the order of the previous points.  This CTA
  // only compacts {previous point, left feature, right feature, distance} of every track into `final` (and the indices of
  // the lost points); the records themselves -- projections, triangulation, 120 bytes per track -- are written by
  // track_emit_kernel on the whole device (in this single CTA they cost 10 us per 1000 tracks, measured with clock64).
  int n_tracks = 0, n_lost = 0, landmarks = 0, accumulated = 0;
  for (int u0 = 0; u0 < n_previous; u0 += kResolveThreads) {
    const int u = u0 + tid;
    int4 t = make_int4(-1, -1, 0, kStatusSkipped);
    if (u < n_previous) t = load_result(tentative + u);
    int total_t, total_l;
    const int pos_t = n_tracks + block_scan_flag(t.w == kStatusTracked, s_red, &total_t);
    const int pos_l = n_lost + block_scan_flag(t.w == kStatusLost, s_red, &total_l);
    if (t.w == kStatusLost) lost[pos_l] = u;
    if (t.w == kStatusTracked) {
      final[pos_t] = make_int4(u, t.x, t.y, t.z);
      atomicAdd(&s_acc[0], previous[u].has_landmark != 0);                                       // :653-655
      atomicAdd(&s_acc[1], t.z);                                                                 // :627
    }
    n_tracks += total_t;
    n_lost += total_l;
  }
  __syncthreads();
  TR_MARK(5)
#ifdef VSLAM_TRACK_TIMING
  if (tid == 0)
    printf("resolve clocks (%d previous, %d + %d features, %d rounds, %d recomputed): reset claims %lld | claim %lld | dirty test %lld | "
           "recompute %lld | exit %lld | gone flags + ordered output %lld\n", n_previous, n_l, n_r, n_rounds, work_total, t_phase[0],
           t_phase[1], t_phase[2], t_phase[3], t_phase[4], t_phase[5]);
#endif
  landmarks = s_acc[0];
  accumulated = s_acc[1];
  if (tid == 0) {
    stats[0] = n_tracks;
    stats[1] = n_lost;
    stats[2] = landmarks;
    stats[3] = accumulated;
  }
}

// K10b: the records of the tracks, one thread per track in its final position (`final` of track_resolve_kernel):
// TrackRecord (:623-643: features, descriptor distance, projections, triangulated camera coordinates) and the bin
// pre-load record compute() reads (:147-155).
__global__ void __launch_bounds__(128) track_emit_kernel(Geometry g, StereoParams sp, TrackParams tp, const int32_t* row_ptr,
                                                         const uint32_t* kp_xy, const uint8_t* desc, const int32_t* n_desc,
                                                         const PreviousPoint* __restrict__ previous,
                                                         const int4* __restrict__ final, const int32_t* __restrict__ stats,
                                                         TrackRecord* __restrict__ tracks, TrackedPoint* __restrict__ tracked,
                                                         const FrameStepState* __restrict__ step) {
  const int k = blockIdx.x * 128 + threadIdx.x;
  if (k >= stats[0]) return;
  if (step) {   // fused frame: see track_search_kernel
#pragma unroll
    for (int i = 0; i < 12; ++i) tp.T[i] = step->T_prior[i];
  }
  const TrackView v = make_view(g, sp, row_ptr, kp_xy, desc, n_desc, nullptr, nullptr);
  const int4 f = final[k];
  const int u = f.x;
  const PreviousPoint* pp = previous + u;
  // the projections of this point (visualisation fields :620-626): same arithmetic as track_point
  double pc[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    pc[i] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(tp.T[4 * i], pp->camera[0]), __dmul_rn(tp.T[4 * i + 1], pp->camera[1])),
                                __dmul_rn(tp.T[4 * i + 2], pp->camera[2])), tp.T[4 * i + 3]);
  const double il0 = __dadd_rn(__dmul_rn(v.fx, pc[0]), __dmul_rn(v.cx, pc[2]));
  const double il1 = __dadd_rn(__dmul_rn(v.fy, pc[1]), __dmul_rn(v.cy, pc[2]));
  const int32_t col_l = to_i32(__ddiv_rn(il0, pc[2])), row_l = to_i32(__ddiv_rn(il1, pc[2]));
  const uint32_t pl = v.xyl[f.y], pr = v.xyr[f.z];
  const float error_x = __fsub_rn((float)col_l, (float)(pl & 0xffffu));
  const float error_y = __fsub_rn((float)row_l, (float)(pl >> 16));
  const double rx = __ddiv_rn(__dadd_rn(il0, v.bx), __dadd_rn(pc[2], 0.0));
  const double ry = __ddiv_rn(__dadd_rn(il1, 0.0), __dadd_rn(pc[2], 0.0));
  TrackRecord r;
  r.index_previous = u;
  r.index_left = f.y;
  r.index_right = f.z;
  r.xl = (float)(pl & 0xffffu);
  r.yl = (float)(pl >> 16);
  r.xr = (float)(pr & 0xffffu);
  r.yr = (float)(pr >> 16);
  r.distance = f.w;
  r.epipolar_offset = (int)(pr >> 16) - (int)(pl >> 16);                                     // :616
  r.projection_left[0] = (float)col_l;
  r.projection_left[1] = (float)row_l;
  r.projection_right[0] = (float)rx;
  r.projection_right[1] = (float)ry;
  r.projection_right_corrected[0] = (float)to_i32(__dsub_rn(rx, (double)error_x));
  r.projection_right_corrected[1] = (float)to_i32(__dsub_rn(ry, (double)error_y));
  r.reserved = 0;
  triangulate(sp, r.xl, r.yl, r.xr, r.yr, r.camera);
  tracks[k] = r;
  TrackedPoint q;
  q.row = (int)(pl >> 16);
  q.col = (int)(pl & 0xffffu);
  q.has_previous = 1;
  q.reserved = 0;
  q.disparity = (double)__fsub_rn(r.xl, r.xr);                                               // frame_point.cpp:19
  q.distance = (double)f.w;
  tracked[k] = q;
}

// ---- recoverPoints ---------------------------------------------------------------------------------------------

// R1: projection and geometric gates (:702-766); slot u of xy[0][.] / xy[1][.] receives the rounded projections of
// lost point u (or a dummy interior pixel when it is rejected; flag in the top bit of the LEFT entry's partner array)
__global__ void __launch_bounds__(256) recover_project_kernel(Geometry g, StereoParams sp, RecoverParams rp,
                                                              const PreviousPoint* __restrict__ lost, int n_lost,
                                                              uint32_t* __restrict__ xy, int stride,
                                                              uint8_t* __restrict__ valid, int32_t* __restrict__ n_xy) {
  const int u = blockIdx.x * 256 + threadIdx.x;
  if (u == 0) n_xy[0] = n_xy[1] = n_lost;
  if (u >= n_lost) return;
  const PreviousPoint* pp = lost + u;
  bool ok = pp->has_landmark != 0;                                                               // :704-706
  double pc[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)                                                                    // :716-717
    pc[i] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(rp.W[4 * i], pp->world[0]), __dmul_rn(rp.W[4 * i + 1], pp->world[1])),
                                __dmul_rn(rp.W[4 * i + 2], pp->world[2])), rp.W[4 * i + 3]);
  const double il0 = __dadd_rn(__dmul_rn(sp.fx, pc[0]), __dmul_rn(sp.cx, pc[2]));                // :725-728
  const double il1 = __dadd_rn(__dmul_rn(sp.fy, pc[1]), __dmul_rn(sp.cy, pc[2]));
  const double il2 = pc[2];
  const double ir0 = __dadd_rn(il0, sp.bx), ir1 = __dadd_rn(il1, 0.0), ir2 = __dadd_rn(il2, 0.0);
  if (il2 < rp.min_depth || il2 > rp.max_depth || ir2 < rp.min_depth || ir2 > rp.max_depth) ok = false;   // :731-736
  const float plx = (float)rint(__ddiv_rn(il0, il2)), ply = (float)rint(__ddiv_rn(il1, il2));    // :739-746
  const float prx = (float)rint(__ddiv_rn(ir0, ir2)), pry = (float)rint(__ddiv_rn(ir1, ir2));
  const float border = __fmul_rn(5.0f, pp->keypoint_size);                                       // :749-750
  const float lo = __fadd_rn(border, 1.0f);
  const float hx = __fsub_rn(__fsub_rn((float)g.cols, border), 1.0f), hy = __fsub_rn(__fsub_rn((float)g.rows, border), 1.0f);
  if (!(plx >= lo && plx <= hx && prx >= lo && prx <= hx && ply >= lo && ply <= hy && pry >= lo && pry <= hy))
    ok = false;                                                                                  // :751-766 (NaN fails)
  // the extractor's own border filter on the (2 border + 1)^2 region would drop the keypoint: descriptor.rows == 0 (:790-792)
  if (border < (float)g.border) ok = false;
  uint32_t l = 31u | (31u << 16), r = l;   // any interior pixel: the descriptor of a rejected slot is never read
  if (ok) {
    l = (uint32_t)(int)plx | ((uint32_t)(int)ply << 16);
    r = (uint32_t)(int)prx | ((uint32_t)(int)pry << 16);
  }
  xy[u] = l;
  xy[stride + u] = r;
  valid[u] = ok;
}

// R3: the three descriptor gates, disparity gate, triangulation and ordered output (:798-858)
__global__ void __launch_bounds__(kResolveThreads) recover_finish_kernel(
    StereoParams sp, RecoverParams rp, const int32_t* n_desc,
    const PreviousPoint* __restrict__ lost, int n_lost, const uint32_t* __restrict__ xy, int stride,
    const uint8_t* __restrict__ valid, const uint8_t* __restrict__ desc, RecoveredRecord* __restrict__ out,
    int32_t* __restrict__ n_out) {
  __shared__ int s_red[32];
  const int tid = threadIdx.x;
  const double thr_tri = triangulation_threshold(sp, n_desc[0]);   // _current_maximum_descriptor_distance_triangulation
  int n = 0;
  for (int u0 = 0; u0 < n_lost; u0 += kResolveThreads) {
    const int u = u0 + tid;
    bool ok = u < n_lost && valid[u];
    uint4 l0, l1, r0, r1;
    int distance = 0;
    uint32_t pl = 0, pr = 0;
    if (ok) {
      const PreviousPoint* pp = lost + u;
      const uint4* dl = reinterpret_cast<const uint4*>(desc + (size_t)u * kDescBytes);
      const uint4* dr = reinterpret_cast<const uint4*>(desc + ((size_t)stride + u) * kDescBytes);
      l0 = dl[0]; l1 = dl[1]; r0 = dr[0]; r1 = dr[1];
      const uint4* ql = reinterpret_cast<const uint4*>(pp->descriptor_left);
      const uint4* qr = reinterpret_cast<const uint4*>(pp->descriptor_right);
      pl = xy[u];
      pr = xy[stride + u];
      if ((double)popc256(ql[0], ql[1], l0, l1) > rp.max_distance_tracking) ok = false;          // :798-802
      if ((double)__fsub_rn((float)(pl & 0xffffu), (float)(pr & 0xffffu)) < sp.min_disparity) ok = false;   // :825-828
      if ((double)popc256(qr[0], qr[1], r0, r1) > rp.max_distance_tracking) ok = false;          // :831-835
      distance = popc256(l0, l1, r0, r1);                                                        // :838-843
      if ((double)distance > thr_tri) ok = false;
    }
    int total;
    const int pos = n + block_scan_flag(ok, s_red, &total);
    if (ok) {
      RecoveredRecord r;
      r.index_lost = u;
      r.distance = distance;
      r.xl = (float)(pl & 0xffffu);
      r.yl = (float)(pl >> 16);
      r.xr = (float)(pr & 0xffffu);
      r.yr = (float)(pr >> 16);
      triangulate(sp, r.xl, r.yl, r.xr, r.yr, r.camera);                                         // :851-855
      *reinterpret_cast<uint4*>(r.descriptor_left) = l0;
      *reinterpret_cast<uint4*>(r.descriptor_left + 16) = l1;
      *reinterpret_cast<uint4*>(r.descriptor_right) = r0;
      *reinterpret_cast<uint4*>(r.descriptor_right + 16) = r1;
      out[pos] = r;
    }
    n += total;
  }
  if (tid == 0) *n_out = n;
}

}  // namespace

void launch_track(const Geometry& g, const StereoParams& sp, const Buffers& b, int pair, const PreviousPoint* previous,
                  int n_previous, const TrackParams& tp, const TrackScratch& s, TrackRecord* tracks, int32_t* lost,
                  TrackedPoint* tracked, cudaStream_t stream, const FrameStepState* step) {
  const int32_t* row_ptr = b.row_ptr + (size_t)2 * pair * (g.rows + 1);
  const uint32_t* kp_xy = b.kp_xy + (size_t)2 * pair * g.cap;
  const uint8_t* desc = b.desc + (size_t)2 * pair * g.cap * kDescBytes;
  const int32_t* n_desc = b.n_desc + 2 * pair;
  uint8_t* gone_l = b.pruned_l + (size_t)pair * g.cap;
  uint8_t* gone_r = b.consumed_r + (size_t)pair * g.cap;
  if (n_previous > 0)
    track_search_kernel<<<(n_previous + 7) / 8, 256, 0, stream>>>(g, sp, tp, row_ptr, kp_xy, desc, n_desc, gone_l, gone_r,
                                                                  previous, n_previous, s.tentative, step);
  // claims, tentative results and worklist in shared memory as far as they fit (claims first)
  // (the opt-in to more than 48 KB of dynamic shared memory is a per-device attribute of the kernel)
  static int limit_of_device[64] = {};   // 0 = not asked yet, -1 = refused
  int device = 0;
  cudaGetDevice(&device);
  int& cached = limit_of_device[device & 63];
  if (cached == 0) {
    int optin = 0;
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    const int limit = std::max(0, optin - 1024);   // (the kernel's static shared memory)
    if (limit > 0 && cudaFuncSetAttribute(track_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, limit) == cudaSuccess) {
      cached = limit;
    } else {
      cudaGetLastError();
      cached = -1;
    }
  }
  const int smem_limit = std::max(cached, 0);
  size_t smem = 0;
  int smem_claims = 0, smem_points = 0;
  const size_t claims_bytes = (sizeof(int32_t) * 2 * (size_t)g.cap + 15) & ~(size_t)15;   // (the int4 results behind the
                                                                                         // claims stay 16-byte aligned)
  if (claims_bytes <= (size_t)smem_limit) {
    smem_claims = g.cap;
    smem = claims_bytes;
    const size_t points = (sizeof(int4) + sizeof(int32_t)) * (size_t)n_previous;
    if (n_previous > 0 && smem + points <= (size_t)smem_limit) {
      smem_points = n_previous;
      smem += points;
    }
  }
  track_resolve_kernel<<<1, kResolveThreads, smem, stream>>>(g, sp, tp, row_ptr, kp_xy, desc, n_desc, gone_l, gone_r,
                                                             previous, n_previous, s.tentative, s.claim_l, s.claim_r,
                                                             s.final, lost, reinterpret_cast<uint8_t*>(tracked), s.stats,
                                                             step, smem_claims, smem_points);
  // (`tracked` doubles as the resolve kernel's per-point dirty flags: one byte per previous point of a 32-byte record)
  if (n_previous > 0)
    track_emit_kernel<<<(n_previous + 127) / 128, 128, 0, stream>>>(g, sp, tp, row_ptr, kp_xy, desc, n_desc, previous, s.final,
                                                                    s.stats, tracks, tracked, step);
}

void launch_recover(const Geometry& g, const StereoParams& sp, const Buffers& b, int pair, const uint8_t* blurred,
                    const PreviousPoint* lost, int n_lost, const RecoverParams& rp, uint32_t* xy, int32_t* n_xy,
                    uint8_t* desc, RecoveredRecord* out, int32_t* n_out, cudaStream_t stream) {
  // scratch layout: xy[2][stride] projections, desc[2][stride][32], valid flags behind the descriptors
  const int stride = n_lost;
  uint8_t* valid = desc + (size_t)2 * stride * kDescBytes;
  if (n_lost > 0) {
    recover_project_kernel<<<(n_lost + 255) / 256, 256, 0, stream>>>(g, sp, rp, lost, n_lost, xy, stride, valid, n_xy);
    if (rp.brief_tests)
      launch_describe_brief(g, reinterpret_cast<const uint16_t*>(blurred), rp.brief_tests, xy, n_xy, desc, stride, 2, stream);
    else
      launch_describe_at(g, blurred, xy, n_xy, desc, stride, 2, stream);
  }
  recover_finish_kernel<<<1, kResolveThreads, 0, stream>>>(sp, rp, b.n_desc + 2 * pair, lost, n_lost, xy, stride,
                                                           valid, desc, out, n_out);
}

// PoseTracker3D::_prunePoints (reference src/position_tracking/pose_tracker_3d.cpp:437-472) on the bin pre-load records of
// the last track(): record k belongs to correspondence k of the aligner; the kept records are compacted in order (their
// position is what compute() reports for a pre-loaded point), in place -- every thread reads its contiguous share
// before anyone writes.  keep = inlier (average error below the kernel) or error != -1 && error < 100 kernel.
namespace {
constexpr int kPruneThreads = 1024;
__global__ void __launch_bounds__(kPruneThreads) prune_tracked_kernel(TrackedPoint* __restrict__ tracked, int n,
                                                                      const double* __restrict__ errors,
                                                                      const uint8_t* __restrict__ inliers,
                                                                      int inliers_only, double error_cap) {
  __shared__ int s_warp[kPruneThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kPer = 8;                       // up to 8192 tracks
  const int begin = tid * kPer;
  TrackedPoint mine[kPer];
  int kept = 0;
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    const int k = begin + j;
    if (k < n) {
      const bool keep = inliers_only ? inliers[k] != 0 : (errors[k] != -1.0 && errors[k] < error_cap);
      if (keep) mine[kept++] = tracked[k];
    }
  }
  int inc = kept;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();                              // (also: every record has been read)
  int base = 0;
  for (int w = 0; w < warp; ++w) base += s_warp[w];
  const int first = base + inc - kept;
  for (int j = 0; j < kept; ++j) tracked[first + j] = mine[j];
}
}  // namespace

void launch_prune_tracked(TrackedPoint* tracked, int n, const double* errors, const uint8_t* inliers, int inliers_only,
                          double error_cap, cudaStream_t stream) {
  if (n <= 0) return;
  prune_tracked_kernel<<<1, kPruneThreads, 0, stream>>>(tracked, n, errors, inliers, inliers_only, error_cap);
}

}  // namespace vslam
