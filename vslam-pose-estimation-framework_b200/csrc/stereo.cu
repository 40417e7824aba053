// stereo.cu -- K5 match_kernel (epipolar row scan, popc-256 Hamming, warp argmin, monotone cursor),
//              K6 select_kernel (bin regularisation + triangulation + ordered output), emit_matches_kernel.
//
// Replaces the scan loop, the bin rule and getPointInLeftCamera of StereoFramePointGenerator::compute
// (reference src/framepoint_generation/stereo_framepoint_generator.cpp:278-426, 147-155/371-394/435-455, 871-895)
// and the adaptive triangulation distance of ::initialize (:109-125).  Semantics: SURVEY.md Appendix A.4/A.5.
//
// Parallel decomposition that keeps the reference's sequential semantics:
//   * within one epipolar pass, image rows are independent (the right cursor never crosses a row except forward
//     to the row's first right feature) -> one warp per (pair, left row); inside the row the left features are
//     visited in ascending column and the cursor `index_R = index_best_R + 1` is carried sequentially;
//   * passes (epipolar offsets 0,+1,-1,...) are separate launches; features matched in an earlier pass are skipped,
//     which equals the reference's prune() because pruning preserves order;
//   * the bin rule is order dependent (partial-order dominance): one thread per bin replays that bin's candidates in
//     emission order (pass, row, col).
#include "kernels.cuh"
#include "stereo_device.cuh"

namespace vslam {

namespace {

__global__ void __launch_bounds__(256) match_kernel(Geometry g, StereoParams sp, const int32_t* __restrict__ row_ptr,
                                                    const uint32_t* __restrict__ kp_xy,
                                                    const uint8_t* __restrict__ desc,
                                                    const int32_t* __restrict__ n_desc, int2* __restrict__ match,
                                                    uint8_t* __restrict__ pruned_l, uint8_t* __restrict__ consumed_r,
                                                    int pass, int offset) {
  const int pair = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= g.rows) return;
  const int rrow = row - offset;   // left.row == right.row + offset
  if (rrow < 0 || rrow >= g.rows) return;
  const int il = 2 * pair, ir = 2 * pair + 1;
  const int32_t* rpl = row_ptr + (size_t)il * (g.rows + 1);
  const int32_t* rpr = row_ptr + (size_t)ir * (g.rows + 1);
  const int lb = rpl[row], le = rpl[row + 1];
  if (lb == le) return;
  const int rb = rpr[rrow], re = rpr[rrow + 1];
  if (rb == re) return;

  const double thr = triangulation_threshold(sp, n_desc[il]);
  const uint32_t* xyl = kp_xy + (size_t)il * g.cap;
  const uint32_t* xyr = kp_xy + (size_t)ir * g.cap;
  const uint4* dl = reinterpret_cast<const uint4*>(desc + (size_t)il * g.cap * kDescBytes);
  const uint4* dr = reinterpret_cast<const uint4*>(desc + (size_t)ir * g.cap * kDescBytes);
  int2* m = match + (size_t)pair * g.cap;
  uint8_t* gone = pruned_l + (size_t)pair * g.cap;
  uint8_t* used = consumed_r + (size_t)pair * g.cap;

  const int n_l = le - lb, n_r = re - rb;
  if (n_l <= 32 && n_r <= 32) {
    // Fast path (a row rarely holds more than 32 features): lane i keeps left feature lb+i, lane s keeps right feature
    // rb+s -- column, pruned flag and the 256-bit descriptor in registers, loaded once with coalesced accesses.  The
    // sequential walk over the left features then only shuffles: no dependent global load per feature.
    int col_mine_l = 0, col_mine_r = 0x7fffffff;
    bool gone_mine = true, used_mine = true;
    uint4 l0 = make_uint4(0, 0, 0, 0), l1 = l0, r0 = l0, r1 = l0;
    if (lane < n_l) {
      col_mine_l = (int)(xyl[lb + lane] & 0xffffu);
      gone_mine = gone[lb + lane] != 0;
      l0 = dl[2 * (lb + lane)];
      l1 = dl[2 * (lb + lane) + 1];
    }
    if (lane < n_r) {
      col_mine_r = (int)(xyr[rb + lane] & 0xffffu);
      used_mine = used[rb + lane] != 0;
      r0 = dr[2 * (rb + lane)];
      r1 = dr[2 * (rb + lane) + 1];
    }
    const unsigned live_l = __ballot_sync(0xffffffffu, !gone_mine);
    int cursor = 0;
    for (unsigned todo = live_l; todo;) {
      const int i = __ffs(todo) - 1;
      todo &= todo - 1;
      const int col_l = __shfl_sync(0xffffffffu, col_mine_l, i);
      uint4 a0, a1;
      a0.x = __shfl_sync(0xffffffffu, l0.x, i); a0.y = __shfl_sync(0xffffffffu, l0.y, i);
      a0.z = __shfl_sync(0xffffffffu, l0.z, i); a0.w = __shfl_sync(0xffffffffu, l0.w, i);
      a1.x = __shfl_sync(0xffffffffu, l1.x, i); a1.y = __shfl_sync(0xffffffffu, l1.y, i);
      a1.z = __shfl_sync(0xffffffffu, l1.z, i); a1.w = __shfl_sync(0xffffffffu, l1.w, i);
      unsigned best = 0xffffffffu;
      if (lane >= cursor && !used_mine && col_l - col_mine_r >= 0)             // :330-335 candidates of this scan
        best = ((unsigned)popc256(a0, a1, r0, r1) << 16) | (unsigned)lane;
      best = __reduce_min_sync(0xffffffffu, best);                             // first strict minimum (:342)
      if (best == 0xffffffffu) continue;
      const int d = (int)(best >> 16), s = (int)(best & 0xffffu);
      if (!((double)d < thr)) continue;                                        // :353
      const int col_r = __shfl_sync(0xffffffffu, col_mine_r, s);
      if ((double)(col_l - col_r) < sp.min_disparity) continue;                // :358-361, cursor NOT advanced
      if (lane == 0) {
        m[lb + i] = make_int2(rb + s, d | (pass << 16));
        gone[lb + i] = 1;
        used[rb + s] = 1;
      }
      cursor = s + 1;                                                          // :414
    }
    return;
  }

  int cursor = rb;
  for (int i = lb; i < le; ++i) {
    if (gone[i]) continue;   // pruned: consumed by track() or matched in an earlier pass
    const int col_l = (int)(xyl[i] & 0xffffu);
    const uint4 a0 = dl[2 * i], a1 = dl[2 * i + 1];
    unsigned best = 0xffffffffu;
    for (int s = cursor + lane; s < re; s += 32) {
      const int col_r = (int)(xyr[s] & 0xffffu);
      if (col_l - col_r < 0) break;            // :333 (columns ascend, so the lane's later candidates fail too)
      if (used[s]) continue;
      const int d = popc256(a0, a1, dr[2 * s], dr[2 * s + 1]);
      const unsigned key = ((unsigned)d << 16) | (unsigned)(s - rb);
      best = min(best, key);                   // strict '<' of :342 == lowest index among equal distances
    }
    best = __reduce_min_sync(0xffffffffu, best);
    if (best == 0xffffffffu) continue;
    const int d = (int)(best >> 16);
    const int s = rb + (int)(best & 0xffffu);
    if (!((double)d < thr)) continue;                                        // :342/:353
    const int col_r = (int)(xyr[s] & 0xffffu);
    if ((double)(col_l - col_r) < sp.min_disparity) continue;                // :358-361, cursor NOT advanced
    if (lane == 0) {
      m[i] = make_int2(s, d | (pass << 16));
      gone[i] = 1;
      used[s] = 1;
    }
    cursor = s + 1;                                                          // :414
  }
}

__device__ __forceinline__ void write_record(const StereoParams& sp, FramePointRecord* o, int i, int s, int dist,
                                             int offset, uint32_t ql, uint32_t qr) {
  FramePointRecord r;
  r.index_left = i;
  r.index_right = s;
  r.xl = (float)(ql & 0xffffu);
  r.yl = (float)(ql >> 16);
  r.xr = (float)(qr & 0xffffu);
  r.yr = (float)(qr >> 16);
  r.distance = dist;
  r.epipolar_offset = offset;
  triangulate(sp, r.xl, r.yl, r.xr, r.yr, r.camera);
  *o = r;
}

__device__ __forceinline__ int pass_to_offset(int pass) {   // :45-50 : 0, +1, -1, +2, -2, ...
  return pass == 0 ? 0 : ((pass & 1) ? (pass + 1) / 2 : -(pass / 2));
}

__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  int woff = 0;
  total = 0;
  for (int w = 0; w < 8; ++w) {
    if (w < warp) woff += s_warp[w];
    total += s_warp[w];
  }
  __syncthreads();
  return woff + inc - v;
}

// K6, generic path (any bin count, arbitrary tracked disparities / distances in double): one CTA per pair, one thread
// per bin replays the bin's candidates in emission order
__global__ void __launch_bounds__(256) select_kernel(Geometry g, StereoParams sp, const int32_t* __restrict__ row_ptr,
                                                     const uint32_t* __restrict__ kp_xy,
                                                     const int32_t* __restrict__ n_desc,
                                                     const int2* __restrict__ match, int n_passes,
                                                     const TrackedPoint* __restrict__ tracked, int n_tracked,
                                                     FramePointRecord* __restrict__ out, int out_cap,
                                                     int32_t* __restrict__ n_out, int32_t* __restrict__ error_flag,
                                                     const int32_t* __restrict__ n_tracked_device) {
  __shared__ int s_warp[8];
  if (n_tracked_device) n_tracked = *n_tracked_device;   // fused frame: the count of the device-side _prunePoints
  const int pair = blockIdx.x;
  const int il = 2 * pair, ir = 2 * pair + 1;
  const int32_t* rpl = row_ptr + (size_t)il * (g.rows + 1);
  const uint32_t* xyl = kp_xy + (size_t)il * g.cap;
  const uint32_t* xyr = kp_xy + (size_t)ir * g.cap;
  const int2* m = match + (size_t)pair * g.cap;
  FramePointRecord* o = out + (size_t)pair * out_cap;
  const int n_bins = g.rows_bin * g.cols_bin;
  const double bs = (double)g.bin_size;

  // number_of_new_points (:397-398)
  {
    const int nl = n_desc[il];
    int c = 0;
    for (int i = threadIdx.x; i < nl; i += 256) c += m[i].x >= 0;
    int total;
    block_exclusive_scan(c, s_warp, total);
    if (threadIdx.x == 0) n_out[2 * pair + 1] = total;
  }

  int carry = 0;
  for (int b0 = 0; b0 < n_bins; b0 += 256) {
    const int bin = b0 + threadIdx.x;
    int win = INT32_MIN;   // >= 0: sorted left index ; < 0 (not MIN): -(k+1) tracked point k
    if (bin < n_bins) {
      const int rbin = bin / g.cols_bin, cbin = bin - rbin * g.cols_bin;
      bool has_prev = false;
      double cur_disp = 0, cur_dist = 0;
      for (int k = 0; k < n_tracked; ++k) {                                   // :147-155
        const TrackedPoint t = tracked[k];
        if (__double2int_rn(__ddiv_rn((double)t.row, bs)) == rbin && __double2int_rn(__ddiv_rn((double)t.col, bs)) == cbin) {
          win = -(k + 1);
          has_prev = t.has_previous != 0;
          cur_disp = t.disparity;
          cur_dist = t.distance;
        }
      }
      const int r_lo = max(0, rbin * g.bin_size - g.bin_size / 2 - 1);
      const int r_hi = min(g.rows - 1, rbin * g.bin_size + g.bin_size / 2 + 1);
      for (int pass = 0; pass < n_passes; ++pass) {
        for (int r = r_lo; r <= r_hi; ++r) {
          if (__double2int_rn(__ddiv_rn((double)r, bs)) != rbin) continue;
          for (int i = rpl[r]; i < rpl[r + 1]; ++i) {
            const int2 mm = m[i];
            if (mm.x < 0 || (mm.y >> 16) != pass) continue;
            const int col = (int)(xyl[i] & 0xffffu);
            if (__double2int_rn(__ddiv_rn((double)col, bs)) != cbin) continue;
            const double disp = (double)(col - (int)(xyr[mm.x] & 0xffffu));    // frame_point.cpp:19
            const double dist = (double)(mm.y & 0xffff);
            if (win != INT32_MIN) {                                           // :378-389
              if (!has_prev && disp > cur_disp && dist <= cur_dist) {
                win = i;
                cur_disp = disp;
                cur_dist = dist;
              }
            } else {                                                          // :390-393
              win = i;
              has_prev = false;
              cur_disp = disp;
              cur_dist = dist;
            }
          }
        }
      }
      if (win != INT32_MIN && has_prev) win = INT32_MIN;                      // :443
    }
    int total;
    const int pos = carry + block_exclusive_scan(win != INT32_MIN, s_warp, total);
    if (win != INT32_MIN) {
      if (pos < out_cap) {
        if (win >= 0) {
          const int2 mm = m[win];
          write_record(sp, &o[pos], win, mm.x, mm.y & 0xffff, pass_to_offset(mm.y >> 16), xyl[win], xyr[mm.x]);
        } else {
          FramePointRecord r = {};
          r.index_left = win;
          r.index_right = -1;
          o[pos] = r;
        }
      } else {
        atomicExch(error_flag, 2);
      }
    }
    carry += total;
  }
  if (threadIdx.x == 0) n_out[2 * pair] = min(carry, out_cap);
}

// rint(v / bin) for non-negative integers, half to even -- exactly what std::rint(static_cast<real>(v)/bin_size)
// yields (:149-152, :372-375): a tie needs 2*rem == bin, otherwise the quotient is >= 1/(2*bin) away from .5
__device__ __forceinline__ int bin_of(int v, int bin) {
  const int q = v / bin, r2 = 2 * (v - q * bin);
  return r2 < bin ? q : (r2 > bin ? q + 1 : q + (q & 1));
}

// K6, fast path.  One CTA (16 warps) per pair, one warp per STRIP of bins (a bin row): the features whose row falls in
// the strip are a contiguous index range of the (row, col)-sorted arrays, so the warp streams them in chunks of 32
// with coalesced loads and replays the matched ones IN ORDER (ballot + shuffle broadcast); the lane that owns bin
// column cb (cb mod 32) updates that bin's state in shared memory.  2.8 k feature visits per pair instead of the
// 160 k of the thread-per-bin kernel above, and the order-dependent rule is still replayed exactly.
// State per bin: winner (sorted left index, or -(k+1) for tracked point k, or INT32_MIN), its disparity and distance
// as float (exact: disparities are differences of integer pixel columns / float-valued inputs, distances Hamming
// counts; the API routes anything else to the generic kernel); distance -1 marks "occupied by a point with
// previous()": `dist_new <= dist_cur` (:385-386) can then never hold, which is the !current->previous() test (:383).
// Replay: the rule is sequential PER BIN only, so the matched features of a chunk of 32 are grouped by bin
// (__match_any_sync) and step j applies the j-th feature of every group at once -- the groups touch different bins, and
// within a group the lane order is the emission order.  Keypoints that are neighbours in (row, col) order rarely share
// a bin, so a chunk takes 1-3 steps instead of one per matched feature.
// kSelectWarps: 16 for batches (one CTA per pair, several CTAs per SM), 32 for a single frame.
template <int kSelectWarps>
__global__ void __launch_bounds__(kSelectWarps * 32) select_strips_kernel(
    Geometry g, StereoParams sp, const int32_t* __restrict__ row_ptr, const uint32_t* __restrict__ kp_xy,
    const int32_t* __restrict__ n_desc, const int2* __restrict__ match, int n_passes,
    const TrackedPoint* __restrict__ tracked, int n_tracked, FramePointRecord* __restrict__ out, int out_cap,
    int32_t* __restrict__ n_out, int32_t* __restrict__ error_flag, const int32_t* __restrict__ n_tracked_device,
    int mode, int32_t* __restrict__ bins) {
  // mode (fused frame, one pair): kSelectAll = everything; kSelectReplay = the replay of the matches over EMPTY bins, the
  // winners go to bins[] (+ the match count in bins[n_bins]) -- it needs neither the aligner nor the prune and runs beside
  // them; kSelectMerge = bins[] minus the bins a surviving track occupies, then the ordered output.  Every pre-loaded point
  // of a fused frame has previous(): no match can take its bin (:383), and it is not appended itself (:443) -- its bin is
  // simply empty in the output, whatever the replay put there.
  extern __shared__ __align__(16) unsigned char s_raw[];
  if (n_tracked_device && mode != kSelectReplay) n_tracked = *n_tracked_device;   // fused frame: the count of the device-side _prunePoints
  const int n_bins = g.rows_bin * g.cols_bin;
  int* s_win = reinterpret_cast<int*>(s_raw);
  float* s_disp = reinterpret_cast<float*>(s_win + n_bins);
  float* s_dist = s_disp + n_bins;
  int* s_cnt = reinterpret_cast<int*>(s_dist + n_bins);   // [rows_bin + 1] winners per strip -> exclusive offsets
  __shared__ int s_matches;

  const int pair = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int il = 2 * pair, ir = 2 * pair + 1;
  const int32_t* rpl = row_ptr + (size_t)il * (g.rows + 1);
  const uint32_t* xyl = kp_xy + (size_t)il * g.cap;
  const uint32_t* xyr = kp_xy + (size_t)ir * g.cap;
  const int2* m = match + (size_t)pair * g.cap;
  FramePointRecord* o = out + (size_t)pair * out_cap;
  const int bs = g.bin_size;

  if (mode == kSelectMerge) {
    for (int i = tid; i < n_bins; i += kSelectWarps * 32) s_win[i] = bins[i];
    if (tid == 0) s_matches = bins[n_bins];
    __syncthreads();
    for (int k = tid; k < n_tracked; k += kSelectWarps * 32) {
      const TrackedPoint t = tracked[k];
      const int trb = bin_of(t.row, bs), tcb = bin_of(t.col, bs);
      if (trb < g.rows_bin && tcb < g.cols_bin) s_win[trb * g.cols_bin + tcb] = INT32_MIN;
    }
    __syncthreads();
  } else {
  for (int i = tid; i < n_bins; i += kSelectWarps * 32) s_win[i] = INT32_MIN;
  if (tid == 0) s_matches = 0;
  __syncthreads();
  if (mode == kSelectReplay) n_tracked = 0;
  // :147-155 pre-load, in order: a later tracked point overwrites an earlier one in the same bin, i.e. the LAST point
  // of a bin wins.  Two parallel passes: atomicMax of the point index per bin (indices are stored as -(k+1) < 0, so the
  // last point is the MINIMUM of the stored values; INT32_MIN = empty is kept apart by the first pass writing through
  // a separate slot), then the winner of each bin writes its disparity / distance.
  for (int k = tid; k < n_tracked; k += kSelectWarps * 32) {
    const TrackedPoint t = tracked[k];
    const int trb = bin_of(t.row, bs), tcb = bin_of(t.col, bs);
    if (trb < g.rows_bin && tcb < g.cols_bin)   // (the reference would write out of bounds)
      atomicMax(reinterpret_cast<unsigned int*>(&s_win[trb * g.cols_bin + tcb]), (unsigned int)(k + 1) | 0x80000000u);
  }
  __syncthreads();
  // s_win now holds 0x80000000 | (k_last + 1) for pre-loaded bins (as unsigned: larger k wins) and INT32_MIN =
  // 0x80000000 for empty ones: decode to -(k_last + 1)
  for (int k = tid; k < n_tracked; k += kSelectWarps * 32) {
    const TrackedPoint t = tracked[k];
    const int trb = bin_of(t.row, bs), tcb = bin_of(t.col, bs);
    if (trb < g.rows_bin && tcb < g.cols_bin) {
      const int b = trb * g.cols_bin + tcb;
      if (((unsigned int)s_win[b] & 0x7fffffffu) == (unsigned int)(k + 1)) {
        s_disp[b] = (float)t.disparity;
        s_dist[b] = t.has_previous ? -1.0f : (float)t.distance;
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < n_bins; i += kSelectWarps * 32) {
    const unsigned int v = (unsigned int)s_win[i];
    if (v != 0x80000000u) s_win[i] = -(int)(v & 0x7fffffffu);
  }
  __syncthreads();
  }

  int my_matches = 0;
  for (int rb = warp; rb < g.rows_bin; rb += kSelectWarps) {
    // rows of this strip: contiguous, within [rb*bs - bs/2 - 1, rb*bs + bs/2 + 1]
    int r_lo = max(0, rb * bs - bs / 2 - 1), r_hi = min(g.rows - 1, rb * bs + bs / 2 + 1);
    while (r_lo <= r_hi && bin_of(r_lo, bs) != rb) ++r_lo;
    while (r_hi >= r_lo && bin_of(r_hi, bs) != rb) --r_hi;
    int winners = 0;
    int* win = s_win + rb * g.cols_bin;
    float* wdist = s_dist + rb * g.cols_bin;
    if (mode != kSelectMerge && r_lo <= r_hi) {
      const int f_lo = rpl[r_lo], f_hi = rpl[r_hi + 1];
      float* wdisp = s_disp + rb * g.cols_bin;
      for (int pass = 0; pass < n_passes; ++pass) {
        // four chunks of 32 features per trip: their (independent) loads are issued together, so a strip pays the
        // global-memory latency of m -> xy_right once per 128 features instead of once per 32
        for (int F0 = f_lo; F0 < f_hi; F0 += 128) {
          int2 mm[4];
          uint32_t ql[4], qr[4];
          bool cand4[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = F0 + 32 * u + lane;
            mm[u] = make_int2(-1, 0);
            ql[u] = 0;
            if (i < f_hi) {
              mm[u] = m[i];
              ql[u] = xyl[i];
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            cand4[u] = mm[u].x >= 0 && (mm[u].y >> 16) == pass;
            qr[u] = cand4[u] ? xyr[mm[u].x] : 0u;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int f0 = F0 + 32 * u;
            const bool cand = cand4[u];          // (all false for chunks past the strip: the replay below is skipped)
            const int col = (int)(ql[u] & 0xffffu);
            const int cb = bin_of(col, bs);
            const float disp = (float)(col - (int)(qr[u] & 0xffffu));   // frame_point.cpp:19
            const float dist = (float)(mm[u].y & 0xffff);
            const unsigned todo = __ballot_sync(0xffffffffu, cand);
            my_matches += __popc(todo);   // cand holds this pass only: the passes sum to every match once
            if (todo) {                   // (uniform)
              // lanes without a match get a key of their own: negative, no bin is
              const unsigned group = __match_any_sync(0xffffffffu, cand ? cb : -1 - lane);
              const int rank = __popc(group & ((1u << lane) - 1u));      // position among the bin's features, in order
              const int steps = __reduce_max_sync(0xffffffffu, cand ? __popc(group) : 0);
              for (int j = 0; j < steps; ++j) {
                if (cand && rank == j) {
                  const int cur = win[cb];
                  if (cur == INT32_MIN || (disp > wdisp[cb] && dist <= wdist[cb])) {   // :390-393 / :378-389
                    win[cb] = f0 + lane;
                    wdisp[cb] = disp;
                    wdist[cb] = dist;
                  }
                }
                __syncwarp();
              }
            }
            __syncwarp();
          }
        }
      }
    }
    // :443 only points without previous() are appended; tracked survivors without previous() are reported
    for (int c0 = 0; c0 < g.cols_bin; c0 += 32) {
      const int c = c0 + lane;
      bool on = false;
      if (c < g.cols_bin) {
        const int w = win[c];
        on = w != INT32_MIN && (mode == kSelectMerge || !(w < 0 && wdist[c] < 0.0f));
        if (!on) win[c] = INT32_MIN;
        if (mode == kSelectReplay) bins[rb * g.cols_bin + c] = on ? w : INT32_MIN;
      }
      winners += __popc(__ballot_sync(0xffffffffu, on));
    }
    if (lane == 0) s_cnt[rb] = winners;
  }
  if (lane == 0 && my_matches) atomicAdd(&s_matches, my_matches);
  __syncthreads();
  if (mode == kSelectReplay) {
    if (tid == 0) bins[n_bins] = s_matches;
    return;
  }
  if (warp == 0) {   // exclusive scan of the strip counts
    int carry = 0;
    for (int r0 = 0; r0 < g.rows_bin; r0 += 32) {
      const int r = r0 + lane;
      const int v = r < g.rows_bin ? s_cnt[r] : 0;
      int inc = v;
      for (int off = 1; off < 32; off <<= 1) {
        const int nb = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += nb;
      }
      if (r < g.rows_bin) s_cnt[r] = carry + inc - v;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) {
      s_cnt[g.rows_bin] = carry;
      n_out[2 * pair] = min(carry, out_cap);
      n_out[2 * pair + 1] = s_matches;
      if (carry > out_cap) atomicExch(error_flag, 2);
    }
  }
  __syncthreads();
  // :435-455 gather the winners row-major over the bin grid; four chunks of 32 bins per trip so that the dependent
  // loads of their winners (match -> xy_right) overlap
  for (int rb = warp; rb < g.rows_bin; rb += kSelectWarps) {
    int pos = s_cnt[rb];
    const int* win = s_win + rb * g.cols_bin;
    for (int C0 = 0; C0 < g.cols_bin; C0 += 128) {
      int w4[4], p4[4];
      int2 mm[4];
      uint32_t ql[4], qr[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = C0 + 32 * u + lane;
        w4[u] = c < g.cols_bin ? win[c] : INT32_MIN;
        const unsigned bal = __ballot_sync(0xffffffffu, w4[u] != INT32_MIN);
        p4[u] = pos + __popc(bal & ((1u << lane) - 1u));
        pos += __popc(bal);
        mm[u] = make_int2(0, 0);
        ql[u] = 0;
        if (w4[u] >= 0) {
          mm[u] = m[w4[u]];
          ql[u] = xyl[w4[u]];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) qr[u] = w4[u] >= 0 ? xyr[mm[u].x] : 0u;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int w = w4[u], p = p4[u];
        if (w != INT32_MIN && p < out_cap) {
          if (w >= 0) {
            write_record(sp, &o[p], w, mm[u].x, mm[u].y & 0xffff, pass_to_offset(mm[u].y >> 16), ql[u], qr[u]);
          } else {
            FramePointRecord r = {};
            r.index_left = w;
            r.index_right = -1;
            o[p] = r;
          }
        }
      }
    }
  }
}

// all new matches of one pair in emission order (pass, row, col): framepoints_new of :163,397, and the output
// of compute() when binning is disabled (:456-460)
__global__ void __launch_bounds__(256) emit_matches_kernel(Geometry g, StereoParams sp,
                                                           const uint32_t* __restrict__ kp_xy,
                                                           const int32_t* __restrict__ n_desc,
                                                           const int2* __restrict__ match, int n_passes,
                                                           FramePointRecord* __restrict__ out, int out_cap,
                                                           int32_t* __restrict__ n_out, int32_t* __restrict__ error_flag) {
  __shared__ int s_warp[8];
  const int pair = blockIdx.x;
  const int il = 2 * pair, ir = 2 * pair + 1;
  const uint32_t* xyl = kp_xy + (size_t)il * g.cap;
  const uint32_t* xyr = kp_xy + (size_t)ir * g.cap;
  const int2* m = match + (size_t)pair * g.cap;
  FramePointRecord* o = out + (size_t)pair * out_cap;
  const int nl = n_desc[il];
  int carry = 0;
  for (int pass = 0; pass < n_passes; ++pass) {
    for (int b0 = 0; b0 < nl; b0 += 256) {
      const int i = b0 + threadIdx.x;
      int2 mm = make_int2(-1, 0);
      if (i < nl) mm = m[i];
      const bool on = mm.x >= 0 && (mm.y >> 16) == pass;
      int total;
      const int pos = carry + block_exclusive_scan(on, s_warp, total);
      if (on) {
        if (pos < out_cap) write_record(sp, &o[pos], i, mm.x, mm.y & 0xffff, pass_to_offset(pass), xyl[i], xyr[mm.x]);
        else atomicExch(error_flag, 2);
      }
      carry += total;
    }
  }
  if (threadIdx.x == 0) n_out[pair] = min(carry, out_cap);
}

}  // namespace

void launch_match(const Geometry& g, const StereoParams& sp, const Buffers& b, int first_pair, int n_pairs, int pass,
                  int epipolar_offset, cudaStream_t stream) {
  if (pass == 0)
    cudaMemsetAsync(b.match + (size_t)first_pair * g.cap, 0xff, sizeof(int2) * (size_t)g.cap * n_pairs, stream);
  dim3 grid((g.rows + 7) / 8, n_pairs);
  match_kernel<<<grid, 256, 0, stream>>>(g, sp, b.row_ptr + (size_t)2 * first_pair * (g.rows + 1),
                                         b.kp_xy + (size_t)2 * first_pair * g.cap,
                                         b.desc + (size_t)2 * first_pair * g.cap * kDescBytes, b.n_desc + 2 * first_pair,
                                         b.match + (size_t)first_pair * g.cap, b.pruned_l + (size_t)first_pair * g.cap,
                                         b.consumed_r + (size_t)first_pair * g.cap, pass, epipolar_offset);
}

bool select_strips_available(const Geometry& g) {
  return (size_t)g.rows_bin * g.cols_bin * 12 + (size_t)(g.rows_bin + 1) * 4 <= 160 * 1024;
}

void launch_select(const Geometry& g, const StereoParams& sp, const Buffers& b, int first_pair, int n_pairs,
                   int n_passes, const TrackedPoint* tracked, int n_tracked, FramePointRecord* out,
                   int out_capacity_per_pair, bool generic, cudaStream_t stream, const int32_t* n_tracked_device,
                   int mode, int32_t* bins) {
  // shared state of the strip kernel: 12 B per bin + one counter per bin row
  const size_t smem = (size_t)g.rows_bin * g.cols_bin * 12 + (size_t)(g.rows_bin + 1) * 4;
  if (!generic && smem <= 160 * 1024) {
    // opt in to > 48 KB dynamic shared memory; the attribute is per device, so it is simply set whenever it is needed
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(n_pairs == 1 ? select_strips_kernel<32> : select_strips_kernel<16>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    auto kernel = n_pairs == 1 ? select_strips_kernel<32> : select_strips_kernel<16>;
    kernel<<<n_pairs, (n_pairs == 1 ? 32 : 16) * 32, smem, stream>>>(
        g, sp, b.row_ptr + (size_t)2 * first_pair * (g.rows + 1), b.kp_xy + (size_t)2 * first_pair * g.cap,
        b.n_desc + 2 * first_pair, b.match + (size_t)first_pair * g.cap, n_passes, tracked, n_tracked,
        out + (size_t)first_pair * out_capacity_per_pair, out_capacity_per_pair, b.n_out + 2 * first_pair, b.error_flag,
        n_tracked_device, mode, bins);
    return;
  }
  select_kernel<<<n_pairs, 256, 0, stream>>>(g, sp, b.row_ptr + (size_t)2 * first_pair * (g.rows + 1),
                                             b.kp_xy + (size_t)2 * first_pair * g.cap, b.n_desc + 2 * first_pair,
                                             b.match + (size_t)first_pair * g.cap, n_passes, tracked, n_tracked,
                                             out + (size_t)first_pair * out_capacity_per_pair, out_capacity_per_pair,
                                             b.n_out + 2 * first_pair, b.error_flag, n_tracked_device);
}

void launch_emit_matches(const Geometry& g, const StereoParams& sp, const Buffers& b, int pair, int n_passes,
                         FramePointRecord* out, int out_capacity, int32_t* n_out, cudaStream_t stream) {
  emit_matches_kernel<<<1, 256, 0, stream>>>(g, sp, b.kp_xy + (size_t)2 * pair * g.cap, b.n_desc + 2 * pair,
                                             b.match + (size_t)pair * g.cap, n_passes, out, out_capacity, n_out,
                                             b.error_flag);
}

}  // namespace vslam
