// fast.cu -- K1 fast_nms_kernel + K2 compact_kernel.
//
// Replaces cv::FastFeatureDetector::detect per detector region as called by
// BaseFramePointGenerator::detectKeypoints (reference src/framepoint_generation/base_framepoint_generator.cpp:
// 23-25, 362-367) and the ordering / border filtering that cv::ORB::compute (:434) and
// IntensityFeatureMatcher::setFeatures + sortFeatureVector (intensity_feature_matcher.cpp:48-79) impose:
// the device-canonical feature order is ascending (row, col), the order the stereo scan consumes.
//
// FAST-9/16 semantics (OpenCV, TYPE_9_16, nonmaxSuppression=true): SURVEY.md Appendix A.1.
#include <cstdlib>

#include <cooperative_groups.h>

#include "kernels.cuh"

namespace vslam {

namespace {

constexpr int TW = 128;            // output tile width  (image-aligned: 128 B = 4 mask words)
constexpr int TH = 54;             // output tile height; TH + 2 = 56 pre-test rows = 7 per warp.  Measured per 4096 KITTI
                                   // pairs: TH 30 -> 7.75 ms, 38 -> 7.37 ms, 54 -> 7.08 ms: the per-tile fixed cost (index
                                   // arithmetic, clears, scan, publish) is paid 1.8x less often and 376 = 7 x 54 - 2 rows
                                   // waste 0.5 % of the tile rows (3.7 % with 30); 7 rows is what one flag register holds
constexpr int RPW = (TH + 2) / 8;  // pre-test rows per warp
constexpr int kSkew = 5;            // word of lane l in pre-test row it: (l + kSkew it) mod 32 (see phase 1)
static_assert((TH + 2) % 8 == 0 && 4 * RPW + 4 <= 32, "one flag register per lane: 4 bits per row + 4 for the halo word");
constexpr int HX = 16;             // smem halo in x (only 4 needed; 16 keeps uint4 loads aligned)
constexpr int SW = TW + 2 * HX + 16;   // 176 bytes = 44 words: consecutive rows are 12 banks apart, so the RPW = 7 consecutive
                                       // rows of a warp start in 7 DIFFERENT banks (0, 12, 24, 4, 16, 28, 8).  With 160-byte
                                       // rows 8 apart, all candidates of a lane shared one bank in the 17 ring look-ups of
                                       // phase 2; the kernel is bound by shared-memory wavefronts (92 % of peak, ncu r1h).
                                       // TMA box rows are multiples of 16 bytes; the 16 extra columns are not read.
constexpr int SH = TH + 8;         // 3 (ring) + 1 (NMS) on both sides
constexpr int CW = TW + 2;         // score tile width (NMS halo 1)
constexpr int CH = TH + 2;
constexpr int CPITCH = 144;        // multiple of 16 (the score tile is cleared with uint4 stores)

// true iff the 16-bit circular mask m has >= 9 contiguous set bits
__device__ __forceinline__ bool arc9(unsigned m) {
  m |= m << 16;
  unsigned a = m & (m >> 1);
  a &= a >> 2;
  a &= a >> 4;
  a &= m >> 8;
  return (a & 0xFFFFu) != 0;
}

// OpenCV cornerScore<16> for a pixel already known to be a corner:
// max(max over 9-arcs of min(d), -(min over 9-arcs of max(d))) - 1 with d[k] = v - ring[k]
__device__ __forceinline__ int corner_score(const int (&d)[16]) {
  int mn[16], mx[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    mn[i] = min(d[i], d[(i + 1) & 15]);
    mx[i] = max(d[i], d[(i + 1) & 15]);
  }
  int mn4[16], mx4[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    mn4[i] = min(mn[i], mn[(i + 2) & 15]);
    mx4[i] = max(mx[i], mx[(i + 2) & 15]);
  }
  int a = -1000, b = 1000;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int m9 = min(min(mn4[i], mn4[(i + 4) & 15]), d[(i + 8) & 15]);
    const int x9 = max(max(mx4[i], mx4[(i + 4) & 15]), d[(i + 8) & 15]);
    a = max(a, m9);
    b = min(b, x9);
  }
  return max(a, -b) - 1;
}

}  // namespace

// ring offsets (dx, dy), k = 0..15 : (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)(-3,-1)(-3,0)(-3,1)(-2,2)(-1,3)
#define VSLAM_RING(F) \
  F(0, 0, 3) F(1, 1, 3) F(2, 2, 2) F(3, 3, 1) F(4, 3, 0) F(5, 3, -1) F(6, 2, -2) F(7, 1, -3) \
  F(8, 0, -3) F(9, -1, -3) F(10, -2, -2) F(11, -3, -1) F(12, -3, 0) F(13, -3, 1) F(14, -2, 2) F(15, -1, 3)

// FAST score of the pixel at p (row pitch `pitch`); 0 if not a corner at threshold t
template <typename Ptr>
__device__ __forceinline__ int fast_score_at(Ptr p, int pitch, int t) {
  const int v = p[0];
  int d[16];
#define F(k, dx, dy) d[k] = v - (int)p[(dy) * pitch + (dx)];
  VSLAM_RING(F)
#undef F
  unsigned dark = 0, bright = 0;   // dark: ring darker than v - t (d > t); bright: d < -t
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    dark |= (unsigned)(d[k] > t) << k;
    bright |= (unsigned)(d[k] < -t) << k;
  }
  if (!(arc9(dark) || arc9(bright))) return 0;
  return corner_score(d);
}

// max(A, -B) of cornerScore<16> with packed 16-bit lanes: lane lo carries d = v - ring, lane hi carries -d, so one
// VIMNMX3.S16x2 chain yields both "max over 9-arcs of min(d)" (lo) and "max over 9-arcs of min(-d)" = -min-max (hi).
// The pixel is a FAST corner at threshold t iff the result is > t; its OpenCV score is the result - 1.
//
// BIASED = true packs with ONE multiply-add per ring pixel: ring * 0xFFFF + v * (1 - 2^16) = d - (d << 16), whose low
// half is d and whose high half is -d - [d < 0] (the borrow of the low half).  The dark lane is therefore one too
// small exactly for the elements that can form a dark arc (d < 0): an arc minimum m >= 1 in that lane is the true
// minimum - 1, and m <= 0 means the true minimum is <= 1.  For thresholds t >= 1 (every threshold the reference's
// controller can produce; t = 0 takes the exact packing) max(lo, hi + 1) is therefore exact whenever it exceeds t,
// and never exceeds t otherwise.
template <bool BIASED>
__device__ __forceinline__ int arc_strength(const uint8_t* p) {
  const int v = p[0];
  unsigned P[16];
  const unsigned vv = (unsigned)v * 0xFFFF0001u;
#define F(k, dx, dy)                                                        \
  {                                                                         \
    const int r = p[(dy) * SW + (dx)];                                      \
    if (BIASED) P[k] = (unsigned)r * 0xFFFFu + vv;                          \
    else P[k] = __byte_perm((unsigned)(v - r), (unsigned)(r - v), 0x5410);  \
  }
  VSLAM_RING(F)
#undef F
  unsigned m3[16], m9[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) m3[i] = __vimin3_s16x2(P[i], P[(i + 1) & 15], P[(i + 2) & 15]);
#pragma unroll
  for (int i = 0; i < 16; ++i) m9[i] = __vimin3_s16x2(m3[i], m3[(i + 3) & 15], m3[(i + 6) & 15]);
  unsigned g[6];
#pragma unroll
  for (int i = 0; i < 5; ++i) g[i] = __vimax3_s16x2(m9[3 * i], m9[3 * i + 1], m9[3 * i + 2]);
  g[5] = m9[15];
  const unsigned m = __vmaxs2(__vimax3_s16x2(g[0], g[1], g[2]), __vimax3_s16x2(g[3], g[4], g[5]));
  return max((int)(short)(m & 0xffffu), (int)(short)(m >> 16) + (BIASED ? 1 : 0));
}

// |d| > t per byte lane (t <= 127): bit 7 of each byte of the result.  VABSDIFF4 is the one native byte-SIMD op.
__device__ __forceinline__ uint32_t absdiff_gt(uint32_t ring, uint32_t center, uint32_t c127_minus_t) {
  const uint32_t ad = __vabsdiffu4(ring, center);
  return (((ad & 0x7f7f7f7fu) + c127_minus_t) | ad) & 0x80808080u;
}

// K1.  Work-efficient FAST: the 16-pixel ring is only evaluated where a cheap necessary condition holds.  After the
// tile is staged, every warp runs a PRIVATE pipeline over its 4 pre-test rows (no atomics, no block barrier between
// the two expensive phases):
//   phase 1  all pixels, 4 per thread, byte-SIMD: every 9-arc contains two ADJACENT compass points (N,E,S,W), so a
//            corner needs (|dN|>t or |dS|>t) and (|dE|>t or |dW|>t)  (~14 % pass); each lane keeps the flags of its 24
//            pixels in one register and one warp scan turns them into the warp's candidate list
//   phase 2  candidates: packed arc minima -> corner decision AND cornerScore<16> -> score tile; the list is
//            compacted in place to the corners inside the tile
//   phase 3  (after one block barrier) strict 3x3 non-maximum suppression of the listed corners -> keypoint bit mask
constexpr int kListCap = RPW * CW + 8;   // candidates of one warp: RPW rows x 130 columns

// compass pre-test of the four pixels of word `wi` in pre-test row `sy`: bit 7 of byte b = pixel b passes
__device__ __forceinline__ uint32_t compass_pretest(const uint8_t (*s_img)[SW], int sy, int wi, uint32_t cadd) {
  const uint32_t* rc = reinterpret_cast<const uint32_t*>(&s_img[sy + 3][0]) + 3 + wi;
  const uint32_t c = rc[0];
  const uint32_t rn = rc[-3 * (SW / 4)];                   // y - 3
  const uint32_t rs = rc[3 * (SW / 4)];                    // y + 3
  const uint32_t re = __funnelshift_r(c, rc[1], 24);       // x + 3
  const uint32_t rw = __funnelshift_r(rc[-1], c, 8);       // x - 3
  return (absdiff_gt(rn, c, cadd) | absdiff_gt(rs, c, cadd)) & (absdiff_gt(re, c, cadd) | absdiff_gt(rw, c, cadd));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256) fast_nms_kernel(const __grid_constant__ CUtensorMap image_map, Geometry g,
                                                       RegionTable rt, int first_image, uint32_t* __restrict__ mask,
                                                       int32_t* __restrict__ raw_count, int single_region,
                                                       const int32_t* __restrict__ thresholds) {
  __shared__ __align__(128) uint8_t s_img[SH][SW];
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ __align__(16) uint8_t s_score[CH][CPITCH];
  __shared__ uint16_t s_list[8][kListCap];     // position codes (sy << 8 | sx)
  __shared__ uint32_t s_mask[TH][4];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int img = single_region ? blockIdx.z : blockIdx.z / g.n_regions;   // (no integer division on the common path)
  const int reg = blockIdx.z - img * g.n_regions;
  const Region R = rt.r[reg];
  // (a captured CUDA graph re-launches this kernel with the same arguments: its thresholds then live in device memory)
  const int t = thresholds ? thresholds[reg] : rt.threshold[reg];
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  // keypoint area of this region, inclusive, image coordinates (FAST skips a 3 px border of its view)
  const int ax0 = R.x + 3, ax1 = R.x + R.w - 4, ay0 = R.y + 3, ay1 = R.y + R.h - 4;
  uint32_t* mrow = mask + (size_t)img * g.rows * g.mask_words;

  const bool hit = !(x0 > ax1 || x0 + TW - 1 < ax0 || y0 > ay1 || y0 + TH - 1 < ay0);
  if (!hit) {
    if (single_region) {  // the mask is written with plain stores: cover this tile with zeros
      for (int i = tid; i < TH * 4; i += 256) {
        const int y = y0 + (i >> 2), wd = (x0 >> 5) + (i & 3);
        if (y < g.rows && wd < g.mask_words) mrow[(size_t)y * g.mask_words + wd] = 0u;
      }
    }
    return;
  }

  // ---- phase 0: ONE TMA box load (cp.async.bulk.tensor.3d, 160 x 38 bytes: the tile + 16 / 4 px halo, zero fill
  // outside the image; x0 - 16 is a multiple of 16 as TMA requires) stages the tile while the threads clear the state
  if (tid == 0) {
    const uint32_t bar = smem_u32(&s_bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(SW * SH) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(&s_img[0][0])), "l"(&image_map), "r"(bar), "r"(x0 - HX), "r"(y0 - 4), "r"(first_image + img)
        : "memory");
  }
  {
    uint4* sc = reinterpret_cast<uint4*>(&s_score[0][0]);
    sc[tid] = make_uint4(0, 0, 0, 0);
    if (tid < CH * CPITCH / 16 - 256) sc[256 + tid] = make_uint4(0, 0, 0, 0);
    if (tid < TH * 4) (&s_mask[0][0])[tid] = 0u;
  }
  __syncthreads();
  {
    const uint32_t bar = smem_u32(&s_bar);
    uint32_t done = 0;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(bar) : "memory");
  }

  // ---- phase 1 (warp-private): compass pre-test on the tile + 1 px NMS halo.  Warp w owns the RPW consecutive
  // pre-test rows from RPW * w.  In row `it` lane l tests the aligned word (l + kSkew it) mod 32 of the tile (image x =
  // x0 + 4 word): SKEWED, because corners cluster along edges -- with one word column per lane a vertical edge put all
  // its candidates into one or two lanes and the per-lane enumeration below ran with 6 of 32 lanes (ncu r2i: 22 % of
  // the kernel's instructions); skewed, a vertical edge is spread over RPW lanes and a horizontal one over all 32, and
  // every row is still read as 32 consecutive words (no bank conflict).  Lanes < 2 RPW also test the two halo words
  // (x0-4.. and x0+128..) of the warp's rows.
  const int cx_lo = max(ax0, x0 - 1), cx_hi = min(ax1, x0 + TW);   // columns whose score is needed
  // the same bounds in tile coordinates (sx = column - (x0 - 1), sy = row - (y0 - 1)); candidates outside them -- they
  // exist only in tiles on the border of a region -- are dropped in phase 2, where every lane is busy, instead of being
  // masked word by word here
  const int sx_lo = cx_lo - (x0 - 1), sx_hi = cx_hi - (x0 - 1);
  const int sy_lo = max(ay0 - (y0 - 1), 0), sy_hi = min(ay1 - (y0 - 1), CH - 1);
  uint16_t* list = s_list[warp];
  const uint32_t lt = (1u << lane) - 1u;
  int ncand = 0;
  if (t <= 127) {
    const uint32_t cadd = 0x01010101u * (uint32_t)(127 - t);
    const int hsy = RPW * warp + min(lane >> 1, RPW - 1), hwi = (lane & 1) ? 33 : 0;   // halo word of lanes < 2 RPW
    // every lane first collects the flags of ITS pixels in one register: bit 8 b + it = byte b of its word in pre-test
    // row it (the pre-test leaves bit 7 of every byte: one shift and one OR per row), bit 8 b + 7 = its halo word ...
    uint32_t flags = 0;
#pragma unroll
    for (int it = 0; it < RPW; ++it)
      flags |= compass_pretest(s_img, RPW * warp + it, ((lane + kSkew * it) & 31) + 1, cadd) >> (7 - it);
    if (lane < 2 * RPW)   // of the left halo word only x0-1 (byte 3) is needed, of the right one only x0+128 (byte 0)
      flags |= compass_pretest(s_img, hsy, hwi, cadd) & (hwi ? 0x00000080u : 0x80000000u);
    // ... then ONE warp scan places the lanes' candidates in the list (the order is irrelevant), instead of four
    // ballots and four predicated stores per pre-test row
    const int mine = __popc(flags);
    int inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += n;
    }
    ncand = __shfl_sync(0xffffffffu, inc, 31);
    uint16_t* dst = list + (inc - mine);
    // entry = lane << 5 | bit of `flags`: the enumeration runs with the few lanes that still hold candidates, so it only
    // stores; phase 2 decodes the entries with all 32 lanes busy
    const uint32_t entry0 = (uint32_t)lane << 5;
    uint32_t left = flags;
    while (left) {
      const uint32_t bit = (uint32_t)__ffs((int)left) - 1u;
      left &= left - 1u;
      *dst++ = (uint16_t)(entry0 + bit);
    }
  } else {   // thresholds above 127 (never produced by the reference's configurations): every pixel is a candidate
    for (int it = 0; it < RPW; ++it) {
      const int sy = RPW * warp + it, iy = y0 - 1 + sy;
      for (int sx0 = 0; sx0 < CW; sx0 += 32) {
        const int sx = sx0 + lane, ix = x0 - 1 + sx;
        const bool on = sx < CW && ix >= cx_lo && ix <= cx_hi && iy >= ay0 && iy <= ay1;
        const unsigned bal = __ballot_sync(0xffffffffu, on);
        if (on) list[ncand + __popc(bal & lt)] = (uint16_t)((sy << 8) + sx);
        ncand += __popc(bal);
      }
    }
  }
  __syncwarp();

  // ---- phase 2 (warp-private): exact segment test + corner score; compact the list in place to in-tile corners
  int ncorner = 0;
  for (int c0 = 0; c0 < ncand; c0 += 32) {
    const int c = c0 + lane;
    int s = 0, code = 0;
    if (c < ncand) {
      code = list[c];
      if (t <= 127) {   // decode lane << 5 | (8 byte + it) -> position code (sy << 8 | sx)
        // entry >> 3 = 4 lane + byte; the tested word was (lane + kSkew it) mod 32, i.e. tile column
        // (4 lane + byte + 4 kSkew it) mod 128, and sx counts from the halo column: + 1
        const int it = code & 7, x4 = code >> 3;
        if (it < RPW) {
          code = ((RPW * warp + it) << 8) + ((x4 + 4 * kSkew * it) & 127) + 1;
        } else {       // halo word of lane (x4 >> 2): left (byte 3 = column x0 - 1) or right (byte 0 = column x0 + 128)
          const int src = x4 >> 2;
          code = ((RPW * warp + min(src >> 1, RPW - 1)) << 8) + ((src & 1) ? CW - 1 : 0);
        }
      }
      const int sy = code >> 8, sx = code & 0xff;
      if (sx >= sx_lo && sx <= sx_hi && sy >= sy_lo && sy <= sy_hi) {   // inside the keypoint area of the region
        const uint8_t* px = &s_img[sy + 3][sx + HX - 1];
        s = t >= 1 ? arc_strength<true>(px) : arc_strength<false>(px);
        if (s > t) s_score[sy][sx] = (uint8_t)(s - 1);
      }
    }
    const int sy = code >> 8, sx = code & 0xff;
    const bool keep = s > t && sy >= 1 && sy <= TH && sx >= 1 && sx <= TW;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    __syncwarp();                                   // every lane has read its entry of this chunk
    if (keep) list[ncorner + __popc(bal & lt)] = (uint16_t)code;
    ncorner += __popc(bal);
  }
  __syncthreads();                                  // the score tile is complete

  // ---- phase 3 (warp-private list): 3x3 strict non-maximum suppression of the corners inside the tile
  for (int c = lane; c < ncorner; c += 32) {
    const int code = list[c];
    const int sy = code >> 8, sx = code & 0xff;
    const int s = s_score[sy][sx];   // (short-circuit on purpose: most corners lose to their first neighbours; a
                                     // branch-free maximum of all 8 measured slower, it is bound by the byte loads)
    const bool kp = s > s_score[sy - 1][sx - 1] && s > s_score[sy - 1][sx] && s > s_score[sy - 1][sx + 1] &&
                    s > s_score[sy][sx - 1] && s > s_score[sy][sx + 1] && s > s_score[sy + 1][sx - 1] &&
                    s > s_score[sy + 1][sx] && s > s_score[sy + 1][sx + 1];
    if (kp) atomicOr(&s_mask[sy - 1][(sx - 1) >> 5], 1u << ((sx - 1) & 31));
  }
  __syncthreads();

  // ---- phase 4: publish the tile's TH x 4 mask words and the raw keypoint count (whole warps enter the shuffles)
  if (tid < ((TH * 4 + 31) & ~31)) {
    const int ry = tid >> 2, wx = tid & 3;
    int found = 0;
    if (ry < TH) {
      const uint32_t word = s_mask[ry][wx];
      const int y = y0 + ry, wd = (x0 >> 5) + wx;
      if (y < g.rows && wd < g.mask_words) {
        if (single_region) mrow[(size_t)y * g.mask_words + wd] = word;
        else if (word) atomicOr(&mrow[(size_t)y * g.mask_words + wd], word);
        found = __popc(word);
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) found += __shfl_xor_sync(0xffffffffu, found, o);
    if (lane == 0 && found) atomicAdd(&raw_count[img * g.n_regions + reg], found);
  }
}

// K2: gridDim.x CTAs per image, each owning a strip of image rows.  Applies the extractor's border filter (31 px
// cv::ORB, 28 px BRIEF-32: KeyPointsFilter::runByImageBorder), builds the CSR row pointer and the (row, col)-sorted
// keypoint list.  One THREAD per image row: a row of the bit mask is a few uint4, all loads of a thread are independent
// (one memory round trip instead of one per row of a warp-per-row loop), the row counts meet in a block scan and every
// thread then emits its own row in order.  A CTA obtains the offset of its strip by re-counting the (L2-resident) mask
// rows above it; batches use one CTA per image, single frames split the image to cut latency.
__global__ void __launch_bounds__(256) compact_kernel(Geometry g, const uint32_t* __restrict__ mask,
                                                      int32_t* __restrict__ row_ptr, uint32_t* __restrict__ kp_xy,
                                                      int32_t* __restrict__ n_desc, int32_t* __restrict__ error_flag,
                                                      uint8_t* __restrict__ pruned_l, uint8_t* __restrict__ consumed_r) {
  extern __shared__ int s_rows[];   // counts of the strip's rows -> exclusive offsets (+1 entry)
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const int img = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  {   // a new frame: no feature is pruned yet (IntensityFeatureMatcher::setFeatures, intensity_feature_matcher.cpp:48-70).
      // The flags of this image (left: pruned, right: consumed) are cleared here instead of by two memsets per chunk.
    uint8_t* flags = ((img & 1) ? consumed_r : pruned_l) + (size_t)(img >> 1) * g.cap;
    const int seg = (g.cap + gridDim.x - 1) / gridDim.x;
    const int end = min(g.cap, (int)(blockIdx.x + 1) * seg);
    for (int i = blockIdx.x * seg + tid; i < end; i += 256) flags[i] = 0;
  }
  const uint32_t* m = mask + (size_t)img * g.rows * g.mask_words;
  const int lo_x = g.border, hi_x = g.cols - g.border;   // keep lo_x <= x < hi_x
  const int lo_y = g.border, hi_y = g.rows - g.border;
  const int strip = (g.rows + gridDim.x - 1) / gridDim.x;
  const int y_begin = blockIdx.x * strip, y_end = min(g.rows, y_begin + strip);
  const bool last = blockIdx.x == gridDim.x - 1;
  // words that can hold a kept column, and the masks of the two partial words
  const int w_lo = lo_x >> 5, w_hi = min(g.mask_words - 1, (hi_x - 1) >> 5);
  const uint32_t m_lo = 0xffffffffu << (lo_x & 31);
  const uint32_t m_hi = ((hi_x & 31) == 0) ? 0xffffffffu : (0xffffffffu >> (32 - (hi_x & 31)));

  auto clip = [&](uint32_t w, int wd) -> uint32_t {
    if (wd < w_lo || wd > w_hi) return 0u;
    if (wd == w_lo) w &= m_lo;
    if (wd == w_hi) w &= m_hi;
    return w;
  };
  auto count_row = [&](int y) -> int {
    if (y < lo_y || y >= hi_y || hi_x <= lo_x) return 0;
    const uint4* row = reinterpret_cast<const uint4*>(m + (size_t)y * g.mask_words);
    int c = 0;
    for (int q = w_lo >> 2; q <= (w_hi >> 2); ++q) {
      const uint4 v = __ldg(row + q);
      c += __popc(clip(v.x, 4 * q)) + __popc(clip(v.y, 4 * q + 1)) + __popc(clip(v.z, 4 * q + 2)) + __popc(clip(v.w, 4 * q + 3));
    }
    return c;
  };

  // offset of the strip: keypoints in the rows above it
  {
    int c = 0;
    for (int y = lo_y + tid; y < min(y_begin, hi_y); y += 256) c += count_row(y);
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) s_warp[warp] = c;
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < 8; ++w) t += s_warp[w];
      s_base = t;
    }
    __syncthreads();
  }

  // per-row counts of the strip
  const int n_rows = y_end - y_begin;
  for (int r = tid; r < n_rows; r += 256) s_rows[r] = count_row(y_begin + r);
  __syncthreads();

  // block-wide exclusive scan over the strip's rows (chunks of 256)
  int carry = s_base;
  for (int b0 = 0; b0 < n_rows; b0 += 256) {
    const int r = b0 + tid;
    const int v = r < n_rows ? s_rows[r] : 0;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += n;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += s_warp[w];
    int total = 0;
    for (int w = 0; w < 8; ++w) total += s_warp[w];
    if (r < n_rows) s_rows[r] = carry + woff + inc - v;
    carry += total;
    __syncthreads();
  }
  if (tid == 0) {
    s_rows[n_rows] = carry;
    if (last) {
      n_desc[img] = min(carry, g.cap);
      if (carry > g.cap) atomicExch(error_flag, 1);
    }
  }
  __syncthreads();
  int32_t* rp = row_ptr + (size_t)img * (g.rows + 1);
  for (int r = tid; r < n_rows + (last ? 1 : 0); r += 256) rp[y_begin + r] = min(s_rows[r], g.cap);

  // ordered emission: every thread writes the keypoints of its rows, ascending column
  uint32_t* xy = kp_xy + (size_t)img * g.cap;
  for (int r = tid; r < n_rows; r += 256) {
    int idx = s_rows[r];
    if (s_rows[r + 1] == idx) continue;
    const int y = y_begin + r;
    const uint4* row = reinterpret_cast<const uint4*>(m + (size_t)y * g.mask_words);
    for (int q = w_lo >> 2; q <= (w_hi >> 2); ++q) {
      const uint4 v = __ldg(row + q);
      const uint32_t w4[4] = {clip(v.x, 4 * q), clip(v.y, 4 * q + 1), clip(v.z, 4 * q + 2), clip(v.w, 4 * q + 3)};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t bits = w4[j];
        while (bits) {
          const int bpos = __ffs(bits) - 1;
          bits &= bits - 1;
          if (idx < g.cap) xy[idx] = (uint32_t)((4 * q + j) * 32 + bpos) | ((uint32_t)y << 16);
          ++idx;
        }
      }
    }
  }
}

// K2 for a single frame: ONE CLUSTER of 8 CTAs per image instead of one thread per row (the batched kernel above keeps
// 2 SMs busy with a frame and is bound by its own instruction latency there: 14 us per KITTI frame, 37 us per 1920 x 1080
// frame, profiles/r2_frame/r2u_prof_frame_step_summary.txt).  The mask of an image is a row-major list of uint4 (mask_words is a multiple of 4); thread t
// of the cluster's 2048 owns the Q consecutive uint4 from t * Q, loads them all at once (independent loads: one memory
// round trip), counts its keypoints; a block scan, the block totals exchanged through distributed shared memory and ONE
// cluster barrier give its offset in the (row, col)-sorted list, and it emits its own bits in order.  The thread that
// owns the first uint4 of a row writes that row's CSR pointer.  Same outputs as compact_kernel.
constexpr int kCompactFrameThreads = 256;
constexpr int kCompactFrameBlocks = 8;   // CTAs per cluster = per image (portable cluster size)
constexpr int kCompactFrameBatch = 8;    // uint4 loaded at once (32 registers)
constexpr int kCompactFrameMaxQ = 64;    // uint4 per thread: images up to 2048 x 64 x 128 = 16.7 M mask bits
__global__ void __launch_bounds__(kCompactFrameThreads) compact_frame_kernel(
    Geometry g, const uint32_t* __restrict__ mask, int32_t* __restrict__ row_ptr, uint32_t* __restrict__ kp_xy,
    int32_t* __restrict__ n_desc, int32_t* __restrict__ error_flag, uint8_t* __restrict__ pruned_l,
    uint8_t* __restrict__ consumed_r, int Q) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ int s_warp[kCompactFrameThreads / 32];
  __shared__ int s_block_total;
  const int img = blockIdx.x / kCompactFrameBlocks;
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ct = rank * kCompactFrameThreads + tid;   // thread index within the image's cluster
  {   // a new frame: no feature is pruned yet (see compact_kernel)
    uint8_t* flags = ((img & 1) ? consumed_r : pruned_l) + (size_t)(img >> 1) * g.cap;
    if ((g.cap & 3) == 0) {   // (the allocations are 256-byte aligned: with a capacity that is a multiple of 4 so is every pair's share)
      uint32_t* words = reinterpret_cast<uint32_t*>(flags);
      for (int i = ct; i < g.cap / 4; i += kCompactFrameThreads * kCompactFrameBlocks) words[i] = 0u;
    } else {
      for (int i = ct; i < g.cap; i += kCompactFrameThreads * kCompactFrameBlocks) flags[i] = 0;
    }
  }
  const uint4* m = reinterpret_cast<const uint4*>(mask + (size_t)img * g.rows * g.mask_words);
  const int per_row = g.mask_words >> 2, total = per_row * g.rows;
  const int lo_x = g.border, hi_x = g.cols - g.border, lo_y = g.border, hi_y = g.rows - g.border;
  const int w_lo = lo_x >> 5, w_hi = min(g.mask_words - 1, (hi_x - 1) >> 5);
  const uint32_t m_lo = 0xffffffffu << (lo_x & 31);
  const uint32_t m_hi = ((hi_x & 31) == 0) ? 0xffffffffu : (0xffffffffu >> (32 - (hi_x & 31)));
  auto clip = [&](uint32_t w, int wd) -> uint32_t {
    if (wd < w_lo || wd > w_hi) return 0u;
    if (wd == w_lo) w &= m_lo;
    if (wd == w_hi) w &= m_hi;
    return w;
  };
  auto clip4 = [&](uint4 v, int y, int xq) -> uint4 {
    if (!(y >= lo_y && y < hi_y && hi_x > lo_x)) return make_uint4(0, 0, 0, 0);
    return make_uint4(clip(v.x, 4 * xq), clip(v.y, 4 * xq + 1), clip(v.z, 4 * xq + 2), clip(v.w, 4 * xq + 3));
  };
  const int q0 = ct * Q;
  const int y0 = q0 / per_row, xq0 = q0 - y0 * per_row;   // row and uint4 column of the thread's first entry

  // pass 1: count
  int count = 0;
  {
    int y = y0, xq = xq0;
    for (int c0 = 0; c0 < Q; c0 += kCompactFrameBatch) {
      uint4 v[kCompactFrameBatch];
#pragma unroll
      for (int i = 0; i < kCompactFrameBatch; ++i)
        v[i] = c0 + i < Q && q0 + c0 + i < total ? __ldg(m + q0 + c0 + i) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int i = 0; i < kCompactFrameBatch; ++i) {
        const uint4 c = clip4(v[i], y, xq);
        count += __popc(c.x) + __popc(c.y) + __popc(c.z) + __popc(c.w);
        if (++xq == per_row) xq = 0, ++y;
      }
    }
  }
  // exclusive scan of the thread counts: within the warp, over the block's warps, over the cluster's blocks
  int inc = count;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  int before = 0, block_total = 0;
#pragma unroll
  for (int w = 0; w < kCompactFrameThreads / 32; ++w) {
    const int c = s_warp[w];
    if (w < warp) before += c;
    block_total += c;
  }
  if (tid == 0) s_block_total = block_total;
  cluster.sync();                           // every block's total is in its shared memory
  int all = 0;
#pragma unroll
  for (int r = 0; r < kCompactFrameBlocks; ++r) {
    const int c = *cluster.map_shared_rank(&s_block_total, r);
    if (r < rank) before += c;
    all += c;
  }
  // pass 2: the same entries again (L1 / L2 hits), emitted in order
  int idx = before + inc - count;
  int32_t* rp = row_ptr + (size_t)img * (g.rows + 1);
  uint32_t* xy = kp_xy + (size_t)img * g.cap;
  {
    int y = y0, xq = xq0;
    for (int c0 = 0; c0 < Q; c0 += kCompactFrameBatch) {
      uint4 v[kCompactFrameBatch];
#pragma unroll
      for (int i = 0; i < kCompactFrameBatch; ++i)
        v[i] = c0 + i < Q && q0 + c0 + i < total ? __ldg(m + q0 + c0 + i) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int i = 0; i < kCompactFrameBatch; ++i) {
        if (c0 + i < Q && q0 + c0 + i < total) {
          if (xq == 0) rp[y] = min(idx, g.cap);
          const uint4 c = clip4(v[i], y, xq);
          const uint32_t w4[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t bits = w4[j];
            while (bits) {
              const int bpos = __ffs(bits) - 1;
              bits &= bits - 1;
              if (idx < g.cap) xy[idx] = (uint32_t)((4 * xq + j) * 32 + bpos) | ((uint32_t)y << 16);
              ++idx;
            }
          }
        }
        if (++xq == per_row) xq = 0, ++y;
      }
    }
  }
  if (ct == 0) {
    rp[g.rows] = min(all, g.cap);
    n_desc[img] = min(all, g.cap);
    if (all > g.cap) atomicExch(error_flag, 1);
  }
  cluster.sync();                           // no block may exit while another still reads its total
}

// one thread per 4 output bytes: aligned 32-bit store, source assembled from two aligned words (rows of the dense
// layout start at arbitrary byte offsets, e.g. stride 1241)
__global__ void __launch_bounds__(256) repitch_kernel(Geometry g, const uint8_t* __restrict__ left,
                                                      const uint8_t* __restrict__ right, int stride,
                                                      uint8_t* __restrict__ image, int32_t* __restrict__ clear,
                                                      int n_clear) {
  // single frames: the raw FAST counters of the frame are zeroed here instead of by a memset node in front of FAST
  if (clear && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
    for (int i = threadIdx.x; i < n_clear; i += 256) clear[i] = 0;
  const int words = g.pitch >> 2;
  const int row = blockIdx.y;
  const int img = blockIdx.z;   // 2 * pair + side
  const uint8_t* src = ((img & 1) ? right : left) + ((size_t)(img >> 1) * g.rows + row) * stride;
  uint32_t* dst = reinterpret_cast<uint32_t*>(image + ((size_t)img * g.rows + row) * g.pitch);
  const size_t a0 = reinterpret_cast<size_t>(src);
  for (int w = blockIdx.x * 256 + threadIdx.x; w < words; w += gridDim.x * 256) {
    uint32_t v = 0;
    if (4 * w < g.cols) {
      const size_t a = a0 + 4 * (size_t)w;
      const uint32_t* p = reinterpret_cast<const uint32_t*>(a & ~(size_t)3);
      const unsigned sh = (unsigned)(a & 3) * 8;
      const uint32_t lo = __ldg(p);
      const uint32_t hi = sh ? __ldg(p + 1) : 0u;   // within the staging slack when it runs past the last row
      v = __funnelshift_r(lo, hi, sh);
      const int left_over = g.cols - 4 * w;         // zero the padding beyond the image width
      if (left_over < 4) v &= 0xffffffffu >> (8 * (4 - left_over));
    }
    dst[w] = v;
  }
}

void launch_repitch(const Geometry& g, const uint8_t* left, const uint8_t* right, int stride, uint8_t* image,
                    int n_pairs, cudaStream_t stream, int32_t* clear, int n_clear) {
  dim3 grid(((g.pitch >> 2) + 255) / 256, g.rows, 2 * n_pairs);
  repitch_kernel<<<grid, 256, 0, stream>>>(g, left, right, stride, image, clear, n_clear);
}

bool make_fast_tensor_map(const Geometry& g, const uint8_t* images, int n_images, CUtensorMap* out) {
  return make_image_tensor_map(g, images, n_images, SW, SH, out);
}

void launch_fast(const Geometry& g, const RegionTable& rt, const Buffers& b, const CUtensorMap& image_map, int first_image,
                 int n_images, cudaStream_t stream, const int32_t* device_thresholds, bool counts_cleared,
                 bool mask_cleared) {
  const int single = g.n_regions == 1;
  const size_t mask_bytes = (size_t)g.rows * g.mask_words * sizeof(uint32_t);
  if (!single && !mask_cleared)
    cudaMemsetAsync(b.mask + (size_t)first_image * g.rows * g.mask_words, 0, mask_bytes * n_images, stream);
  if (!counts_cleared)
    cudaMemsetAsync(b.raw_count + (size_t)first_image * g.n_regions, 0, sizeof(int32_t) * g.n_regions * n_images, stream);
  dim3 grid((g.cols + TW - 1) / TW, (g.rows + TH - 1) / TH, n_images * g.n_regions);
  fast_nms_kernel<<<grid, 256, 0, stream>>>(image_map, g, rt, first_image,
                                            b.mask + (size_t)first_image * g.rows * g.mask_words,
                                            b.raw_count + (size_t)first_image * g.n_regions, single, device_thresholds);
}

void launch_compact(const Geometry& g, const Buffers& b, int first_image, int n_images, cudaStream_t stream) {
  const int total = (g.mask_words >> 2) * g.rows;
  const int cluster_threads = kCompactFrameThreads * kCompactFrameBlocks;
  const int Q = (total + cluster_threads - 1) / cluster_threads;
  if (n_images <= 16 && Q <= kCompactFrameMaxQ) {   // single frames: one cluster of 8 CTAs per image (latency)
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr;
    cfg.gridDim = dim3(n_images * kCompactFrameBlocks);
    cfg.blockDim = dim3(kCompactFrameThreads);
    cfg.stream = stream;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = kCompactFrameBlocks;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, compact_frame_kernel, g, (const uint32_t*)(b.mask + (size_t)first_image * g.rows * g.mask_words),
                           b.row_ptr + (size_t)first_image * (g.rows + 1), b.kp_xy + (size_t)first_image * g.cap,
                           b.n_desc + first_image, b.error_flag, b.pruned_l + (size_t)(first_image >> 1) * g.cap,
                           b.consumed_r + (size_t)(first_image >> 1) * g.cap, Q) == cudaSuccess)
      return;
    cudaGetLastError();   // no cluster launch on this device: the strip kernel below
  }
  const int strips = n_images <= 16 ? 8 : 1;   // single frames: split each image over 8 CTAs (latency)
  const size_t smem = sizeof(int) * ((g.rows + strips - 1) / strips + 2);
  compact_kernel<<<dim3(strips, n_images), 256, smem, stream>>>(
      g, b.mask + (size_t)first_image * g.rows * g.mask_words, b.row_ptr + (size_t)first_image * (g.rows + 1),
      b.kp_xy + (size_t)first_image * g.cap, b.n_desc + first_image, b.error_flag,
      b.pruned_l + (size_t)(first_image >> 1) * g.cap, b.consumed_r + (size_t)(first_image >> 1) * g.cap);
}

// cv::KeyPoint::response of the kept keypoints (cornerScore<16>): nothing on the path reads it, so it is produced
// only when a caller asks for keypoints (vslam_fpg_get_features).  For a corner the score does not depend on the
// detector threshold: max(t, A, -B) - 1 with A > t or -B > t.
__global__ void __launch_bounds__(256) score_kernel(Geometry g, const uint8_t* __restrict__ image,
                                                    const uint32_t* __restrict__ kp_xy,
                                                    const int32_t* __restrict__ n_desc, uint8_t* __restrict__ kp_score) {
  const int n = n_desc[0];
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const uint32_t q = kp_xy[i];
    kp_score[i] = (uint8_t)fast_score_at(image + (size_t)(q >> 16) * g.pitch + (q & 0xffffu), g.pitch, 0);
  }
}

// The features of one stereo pair packed for ONE device-to-host copy (the prefetch that vslam_fpg_initialize starts):
// per side s, at byte offset feature_pack_offset(s): [n] u32 positions | [n] u8 FAST responses, padded to 16 | [n][32]
// descriptors, n = n_desc[s].  gridDim.y = side.
__global__ void __launch_bounds__(256) pack_features_kernel(Geometry g, const uint8_t* __restrict__ image,
                                                            const uint32_t* __restrict__ kp_xy,
                                                            const uint8_t* __restrict__ desc,
                                                            const int32_t* __restrict__ n_desc, uint8_t* __restrict__ out) {
  const int side = blockIdx.y;
  const int n = n_desc[side];
  const size_t base = side == 0 ? 0 : feature_pack_bytes(n_desc[0]);
  uint32_t* o_xy = reinterpret_cast<uint32_t*>(out + base);
  uint8_t* o_score = out + base + sizeof(uint32_t) * (size_t)n;
  uint4* o_desc = reinterpret_cast<uint4*>(out + base + feature_pack_desc_offset(n));
  const uint8_t* img = image + (size_t)side * g.rows * g.pitch;
  const uint32_t* xy = kp_xy + (size_t)side * g.cap;
  const uint4* d = reinterpret_cast<const uint4*>(desc + (size_t)side * g.cap * kDescBytes);
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const uint32_t q = xy[i];
    o_xy[i] = q;
    o_score[i] = (uint8_t)fast_score_at(img + (size_t)(q >> 16) * g.pitch + (q & 0xffffu), g.pitch, 0);
    o_desc[2 * i] = d[2 * i];
    o_desc[2 * i + 1] = d[2 * i + 1];
  }
}

void launch_pack_features(const Geometry& g, const Buffers& b, uint8_t* out, cudaStream_t stream) {
  pack_features_kernel<<<dim3(32, 2), 256, 0, stream>>>(g, b.image, b.kp_xy, b.desc, b.n_desc, out);
}

void launch_score(const Geometry& g, const Buffers& b, int image, cudaStream_t stream) {
  score_kernel<<<32, 256, 0, stream>>>(g, b.image + (size_t)image * g.rows * g.pitch, b.kp_xy + (size_t)image * g.cap,
                                       b.n_desc + image, b.kp_score + (size_t)image * g.cap);
}

}  // namespace vslam
