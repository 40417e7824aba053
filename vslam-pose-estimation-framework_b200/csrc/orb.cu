// orb.cu -- K3 blur_kernel (7x7 sigma 2 Gaussian, float separable, rounded to u8) + K4 describe_kernel (rBRIEF-256).
//
// Replaces cv::ORB::create()->compute(image, keypoints, descriptors) as called by
// BaseFramePointGenerator::computeDescriptors (reference src/framepoint_generation/base_framepoint_generator.cpp:
// 195, 222, 431-438).  Semantics: SURVEY.md Appendix A.3.  The float evaluation order is the contract shared with
// the CPU oracle (oracle/c/vslam_oracle.c, orc_gauss7_u8):
//   row pass   : acc = k[0]*p[-3]; acc = fma(k[i], p[i-3], acc), i = 1..6
//   column pass: acc = k[3]*r[0];  acc = fma(k[3+j], r[+j] + r[-j], acc), j = 1..3 ; out = rint(acc)
#include <utility>

#include "kernels.cuh"

namespace vslam {

namespace {

struct GaussKernel {
  float k[7];
};

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
  return i;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// the same byte of two words -> (float, float), exact and without the (quarter-rate) I2F unit: PRMT places each byte
// in the mantissa of 2^23, ONE packed FADD2 subtracts 2^23 from both
__device__ __forceinline__ float2 bytes_to_float2(uint32_t wa, uint32_t wb, int j) {
  return __fadd2_rn(make_float2(__uint_as_float(__byte_perm(wa, 0x4B000000u, 0x7650u | (unsigned)j)),
                                __uint_as_float(__byte_perm(wb, 0x4B000000u, 0x7650u | (unsigned)j))),
                    make_float2(-8388608.0f, -8388608.0f));
}

// K3.  One WARP per 128 x 48 output tile, no block barrier and no intermediate in shared memory.  Lane 0 stages the
// tile + halo (160 x 54 bytes, zero fill outside the image) with one TMA box load; BORDER_REFLECT_101 is then patched
// into the out-of-image rows / columns of the staged tile.  Every lane owns 4 columns and walks down the tile two rows
// per step with the 7-row window of the column pass in registers:
//   row pass    : rows (2i, 2i+1) together, packed FFMA2 (.x = row 2i, .y = row 2i+1)          -> E_i = (t[2i], t[2i+1])
//   O_i         : (t[2i+1], t[2i+2]) = (E_i.y, E_{i+1}.x), the pairs of the other parity
//   column pass : outputs centred on (t[2j+1], t[2j+2]) = O_j:  k3 O_j + k4 (E_{j+1} + E_j) + k5 (O_{j+1} + O_{j-1})
//                 + k6 (E_{j+2} + E_{j-1}), all packed FADD2 / FFMA2, in the oracle's order
// Packed ops are two IEEE fp32 operations per instruction: every result equals the scalar evaluation bit for bit.
constexpr int BL_W = 128;               // columns of a warp tile (4 per lane)
#ifndef VSLAM_BL_H
#define VSLAM_BL_H 47
#endif
constexpr int BL_H = VSLAM_BL_H;        // output rows of a warp tile (376 = 8 x 47: no KITTI tile row is wasted)
constexpr int BL_SW = BL_W + 32;        // staged columns x0-16 .. x0+143 (TMA rows are multiples of 16 bytes)
constexpr int BL_SH = BL_H + 6;         // staged rows y0-3 .. y0+52
constexpr int BL_WALK = (BL_SH + 7) / 8 * 8;   // the walk is unrolled over 4 row pairs (the period of the register rings):
                                               // it may run up to 7 rows past the staged ones; what it computes from
                                               // them lands in output rows >= BL_H, which are never stored
#ifndef VSLAM_BL_WARPS
#define VSLAM_BL_WARPS 4
#endif
constexpr int BL_WARPS = VSLAM_BL_WARPS;       // independent warp tiles per CTA (stacked in y)
constexpr int BL_TILE = (BL_SW * BL_WALK + 127) / 128 * 128;   // TMA destinations are 128-byte aligned

__global__ void __launch_bounds__(BL_WARPS * 32, 4) blur_kernel(const __grid_constant__ CUtensorMap image_map, Geometry g,
                                                                GaussKernel gk, int first_image,
                                                                uint8_t* __restrict__ blurred) {
  __shared__ __align__(128) uint8_t s_in[BL_WARPS][BL_TILE];
  __shared__ __align__(8) unsigned long long s_bar[BL_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int img = blockIdx.z;
  const int x0 = blockIdx.x * BL_W, y0 = (blockIdx.y * BL_WARPS + warp) * BL_H;
  if (y0 >= g.rows) return;             // warps are independent: no block-wide barrier below
  uint8_t* tile = s_in[warp];
  const int tx0 = x0 - 16, ty0 = y0 - 3;   // image coordinates of the staged tile's origin

  const uint32_t bar = smem_u32(&s_bar[warp]);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(BL_SW * BL_SH) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(tile)), "l"(&image_map), "r"(bar), "r"(tx0), "r"(ty0), "r"(first_image + img)
        : "memory");
  }
  __syncwarp();
  {
    uint32_t done = 0;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(bar) : "memory");
  }

  // ---- BORDER_REFLECT_101: rows first (whole staged rows), then the three columns on either side of the image
  if (ty0 < 0 || ty0 + BL_SH > g.rows) {
    uint32_t* t32 = reinterpret_cast<uint32_t*>(tile);
    for (int r = 0; r < BL_SH; ++r) {
      const int gy = ty0 + r;
      if (gy >= 0 && gy < g.rows) continue;
      const int src = reflect101(gy, g.rows) - ty0;
      if (src < 0 || src >= BL_SH) continue;       // a row no stored output reads
      for (int w = lane; w < BL_SW / 4; w += 32) t32[r * (BL_SW / 4) + w] = t32[src * (BL_SW / 4) + w];
    }
    __syncwarp();
  }
  if (x0 == 0 || x0 + BL_W + 3 > g.cols) {
    for (int i = lane; i < BL_SH * 6; i += 32) {
      const int r = i / 6, j = i - r * 6;
      const int x = j < 3 ? -1 - j : g.cols + (j - 3);
      const int dst = x - tx0, src = reflect101(x, g.cols) - tx0;
      if (dst >= 0 && dst < BL_SW && src >= 0 && src < BL_SW) tile[r * BL_SW + dst] = tile[r * BL_SW + src];
    }
    __syncwarp();
  }

  // ---- the walk.  Lane l owns columns x0 + 4l .. +3: inputs x0 + 4l - 3 .. + 6 lie in the three words from staged
  // column 12 + 4l (word 3 + l: consecutive lanes, consecutive banks).
  const uint32_t* w0 = reinterpret_cast<const uint32_t*>(tile) + 3 + lane;
  float2 kk[7];
#pragma unroll
  for (int t = 0; t < 7; ++t) kk[t] = make_float2(gk.k[t], gk.k[t]);
  const float2 bias = make_float2(12582912.0f, 12582912.0f);   // 1.5 * 2^23: rint(acc) lands in the low mantissa byte
  uint8_t* o = blurred + ((size_t)img * g.rows + y0) * g.pitch + x0 + 4 * lane;
  const int n_y = min(BL_H, g.rows - y0);   // output rows of this tile that exist
  // rings of the last four E and O pairs, indexed by compile-time constants only (registers).  The loop is unrolled
  // over ONE ring period (4 steps): a fully unrolled walk (3.9 k instructions) ran out of the instruction cache
  // (39 % of the warp stalls were "no instruction").
  float2 E[4][4], O[4][4];
#pragma unroll 1
  for (int i4 = 0; i4 < BL_WALK / 2; i4 += 4) {
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4) {
      const int i = i4 + s4;
      const uint32_t* wa = w0 + (2 * i) * (BL_SW / 4);
      const uint32_t* wb = wa + BL_SW / 4;
      const uint32_t a0 = wa[0], a1 = wa[1], a2 = wa[2], b0 = wb[0], b1 = wb[1], b2 = wb[2];
      float2 p[10];   // .x: staged row 2i, .y: staged row 2i + 1; p[j] = input column 4l - 3 + j
      p[0] = bytes_to_float2(a0, b0, 1);
      p[1] = bytes_to_float2(a0, b0, 2);
      p[2] = bytes_to_float2(a0, b0, 3);
#pragma unroll
      for (int j = 0; j < 4; ++j) p[3 + j] = bytes_to_float2(a1, b1, j);
      p[7] = bytes_to_float2(a2, b2, 0);
      p[8] = bytes_to_float2(a2, b2, 1);
      p[9] = bytes_to_float2(a2, b2, 2);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float2 acc = __fmul2_rn(kk[0], p[c]);
#pragma unroll
        for (int t = 1; t < 7; ++t) acc = __ffma2_rn(kk[t], p[c + t], acc);
        O[(s4 + 3) & 3][c] = make_float2(E[(s4 + 3) & 3][c].y, acc.x);   // O_{i-1} (garbage at i = 0, never read)
        E[s4][c] = acc;                                                   // E_i
      }
      if (i >= 3) {   // outputs centred on O_j, j = i - 2: image rows y0 + 2 (j - 1) and + 1
        float2 r[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float2 acc = __fmul2_rn(kk[3], O[(s4 + 2) & 3][c]);                                             // O_j
          acc = __ffma2_rn(kk[4], __fadd2_rn(E[(s4 + 3) & 3][c], E[(s4 + 2) & 3][c]), acc);              // E_{j+1} + E_j
          acc = __ffma2_rn(kk[5], __fadd2_rn(O[(s4 + 3) & 3][c], O[(s4 + 1) & 3][c]), acc);              // O_{j+1} + O_{j-1}
          acc = __ffma2_rn(kk[6], __fadd2_rn(E[s4][c], E[(s4 + 1) & 3][c]), acc);                         // E_{j+2} + E_{j-1}
          // acc is a convex combination (weights sum to 1 within 1e-7) of values in [0, 255]: it cannot leave
          // [0, 255.0001], so cv::saturate_cast's clamp is a no-op and is not evaluated
          r[c] = __fadd2_rn(acc, bias);
        }
        const int k = 2 * (i - 3);
        if (k < n_y)
          *reinterpret_cast<uint32_t*>(o) =
              __byte_perm(__byte_perm(__float_as_uint(r[0].x), __float_as_uint(r[1].x), 0x0040),
                          __byte_perm(__float_as_uint(r[2].x), __float_as_uint(r[3].x), 0x0040), 0x5410);
        if (k + 1 < n_y)
          *reinterpret_cast<uint32_t*>(o + g.pitch) =
              __byte_perm(__byte_perm(__float_as_uint(r[0].y), __float_as_uint(r[1].y), 0x0040),
                          __byte_perm(__float_as_uint(r[2].y), __float_as_uint(r[3].y), 0x0040), 0x5410);
        o += 2 * (size_t)g.pitch;
      }
    }
  }
}

// K4.  One LANE per keypoint.  A warp stages the 26-row x 32-byte patches of its 32 keypoints in shared memory
// (coalesced 32-bit loads from the 4-byte aligned start of each row, 8 words cover x-13..x+12 at any alignment);
// lane l then evaluates all 256 tests of keypoint l from ITS patch.  All lanes read the same pattern offset at the
// same time and the patch stride is an odd number of words, so the 512 byte look-ups per lane are bank-conflict free
// and their offsets are compile-time immediates; each lane stores its 32-byte descriptor as two uint4.
constexpr int PR = 26;                    // patch rows
constexpr int PP = 32;                    // patch row pitch (bytes): 8 aligned words
constexpr int PSTRIDE = PR * PP + 4;      // 836 bytes = 209 words (odd): lane l's patch starts in bank 17*l mod 32
constexpr int DWARPS = 4;                 // warps per CTA
constexpr int DSMEM = DWARPS * 32 * PSTRIDE;

struct OrbPattern {
  int8_t v[1024];
};
constexpr OrbPattern kOrb = {{
#include "orb_pattern_31.inc"
}};

// test K of the pattern as compile-time offsets into a buffer of row pitch PITCH whose origin is the pixel (x-13, y-13)
// of the keypoint (the table is only ever read in constant expressions)
template <int K, int PITCH>
struct OrbTest {
  static constexpr int o0 = (kOrb.v[4 * K + 1] + 13) * PITCH + kOrb.v[4 * K] + 13;
  static constexpr int o1 = (kOrb.v[4 * K + 3] + 13) * PITCH + kOrb.v[4 * K + 2] + 13;
};

// descriptor word W: tests 32W .. 32W+31, bit j of byte i = test 8i + j (LSB first).  The tests are visited from
// j = 31 down to 0 and the sign bit of t0 - t1 (set iff t0 < t1) is shifted in with one funnel shift per test.
template <int W, int PITCH, int... J>
__device__ __forceinline__ uint32_t brief_word_impl(const uint8_t* c, std::integer_sequence<int, J...>) {
  uint32_t w = 0;
  ((w = __funnelshift_l((uint32_t)((int)c[OrbTest<32 * W + 31 - J, PITCH>::o0] - (int)c[OrbTest<32 * W + 31 - J, PITCH>::o1]), w, 1)), ...);
  return w;
}

template <int W, int PITCH = PP>
__device__ __forceinline__ uint32_t brief_word(const uint8_t* c) {
  return brief_word_impl<W, PITCH>(c, std::make_integer_sequence<int, 32>{});
}

// K4, tile form.  The keypoints of one image are (row, col)-sorted with a CSR row pointer, so the keypoints whose centre
// lies in a 224 x 64 tile are a filtered slice of one contiguous index range.  One CTA per tile: a single TMA box load
// (cp.async.bulk.tensor.3d, 256 x 90 bytes = the tile + the 13 px pattern radius, out-of-bounds zero-filled) stages the
// blurred pixels while the threads compact the tile's keypoints into a list; then one lane per keypoint evaluates the
// 256 tests straight from the shared tile (offsets are compile-time immediates).  Every blurred byte is fetched about
// 1.7x (halo) instead of once per overlapping 26 x 32 patch (~5x), and no per-keypoint staging instructions remain.
#ifndef VSLAM_DT_H
#define VSLAM_DT_H 64
#endif
constexpr int DT_W = 224, DT_H = VSLAM_DT_H;    // keypoint-centre area of a tile
constexpr int DT_BW = 256, DT_BH = DT_H + 26;   // TMA box: columns x0-15 .. x0+240, rows y0-13 .. y0+76
constexpr int DT_X = 15;                        // the box starts 15 px left of the tile: x0 - 15 = 16 (1 + 14 k), and TMA
                                                // needs every row of the box to start at a 16-byte aligned address
static_assert(DT_W % 16 == 0 && (31 - DT_X) % 16 == 0 && DT_X >= 13 && DT_W + 12 + DT_X < DT_BW, "tile geometry");
constexpr int DT_THREADS = 128;
constexpr int DT_LIST = 512;                    // keypoints compacted per round
constexpr int DT_BINS = 128;                    // sort key of a keypoint: byte column of its patch origin mod 128 =
                                                // (shared-memory bank, byte in word)
static_assert(DT_BINS == DT_THREADS, "one bin per thread");

__global__ void __launch_bounds__(DT_THREADS) describe_tile_kernel(const __grid_constant__ CUtensorMap blurred_map,
                                                                   Geometry g, const int32_t* __restrict__ row_ptr,
                                                                   const uint32_t* __restrict__ kp_xy,
                                                                   uint8_t* __restrict__ desc, int scratch_first) {
  __shared__ __align__(128) uint8_t s_tile[DT_BH][DT_BW];
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ uint32_t s_q[DT_LIST];
  __shared__ int s_f[DT_LIST];
  __shared__ int s_bin[DT_BINS];
  __shared__ int s_warp[DT_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int img = blockIdx.z;
  const int x0 = 31 + blockIdx.x * DT_W, y0 = 31 + blockIdx.y * DT_H;
  const int y1 = min(y0 + DT_H, g.rows - 31);   // keypoints live in [31, rows - 31) x [31, cols - 31)
  if (y0 >= y1) return;
  const int32_t* rp = row_ptr + (size_t)img * (g.rows + 1);
  const int f0 = rp[y0], f1 = rp[y1];
  if (f0 == f1) return;
  const uint32_t* xy = kp_xy + (size_t)img * g.cap;
  uint4* out = reinterpret_cast<uint4*>(desc + (size_t)img * g.cap * kDescBytes);

  if (tid == 0) {
    const uint32_t bar = smem_u32(&s_bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(DT_BW * DT_BH) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(&s_tile[0][0])), "l"(&blurred_map), "r"(bar), "r"(x0 - DT_X), "r"(y0 - 13), "r"(scratch_first + img)
        : "memory");
  }
  // The 375 byte look-ups of a lane go to the bank of its keypoint's tile column (the tile pitch is 2 x 32 banks, so
  // the row does not matter): 32 lanes with the columns of 32 arbitrary keypoints collide ~3.3-fold, and the shared
  // memory wavefronts, not the instruction issue, bound this kernel.  The keypoints of the tile are therefore counting-
  // sorted by (bank, byte in word) of their patch origin and DEALT to the warp rounds like cards, so that one round
  // holds at most ceil(bin size / rounds) keypoints of a bank (~2.1-fold collisions, simulated and measured).
  bool tile_ready = false;
  for (int base = f0; base < f1; base += DT_LIST) {
    s_bin[tid] = 0;
    __syncthreads();
    const int end = min(f1, base + DT_LIST);
    for (int f = base + tid; f < end; f += DT_THREADS) {
      const int bx = (int)(xy[f] & 0xffffu) - x0;
      if (bx >= 0 && bx < DT_W) atomicAdd(&s_bin[(bx + DT_X - 13) & (DT_BINS - 1)], 1);
    }
    __syncthreads();
    int n;
    {   // exclusive scan of the 128 bins (one per thread)
      const int v = s_bin[tid];
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (lane == 31) s_warp[warp] = inc;
      __syncthreads();
      int off = 0;
      n = 0;
#pragma unroll
      for (int w = 0; w < DT_THREADS / 32; ++w) {
        const int t = s_warp[w];
        if (w < warp) off += t;
        n += t;
      }
      s_bin[tid] = off + inc - v;
    }
    const int rounds = (n + 31) >> 5;   // warp rounds of this chunk; sorted position p -> slot (p % rounds, p / rounds)
    for (int p = n + tid; p < 32 * rounds; p += DT_THREADS) s_q[(p % rounds) * 32 + p / rounds] = 0xffffffffu;
    __syncthreads();
    for (int f = base + tid; f < end; f += DT_THREADS) {
      const uint32_t q = xy[f];
      const int bx = (int)(q & 0xffffu) - x0;
      if (bx >= 0 && bx < DT_W) {
        const int p = atomicAdd(&s_bin[(bx + DT_X - 13) & (DT_BINS - 1)], 1);
        const int slot = (p % rounds) * 32 + p / rounds;
        s_q[slot] = q;
        s_f[slot] = f;
      }
    }
    __syncthreads();
    if (!tile_ready) {   // wait for the TMA box (phase 0 of the barrier)
      const uint32_t bar = smem_u32(&s_bar);
      uint32_t done = 0;
      while (!done)
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
            : "=r"(done) : "r"(bar) : "memory");
      tile_ready = true;
    }
    for (int k = tid; k < 32 * rounds; k += DT_THREADS) {
      const uint32_t q = s_q[k];
      if (q == 0xffffffffu) continue;
      const uint8_t* c = &s_tile[(int)(q >> 16) - y0][(int)(q & 0xffffu) - x0 + (DT_X - 13)];
      uint4 lo, hi;
      lo.x = brief_word<0, DT_BW>(c);
      lo.y = brief_word<1, DT_BW>(c);
      lo.z = brief_word<2, DT_BW>(c);
      lo.w = brief_word<3, DT_BW>(c);
      hi.x = brief_word<4, DT_BW>(c);
      hi.y = brief_word<5, DT_BW>(c);
      hi.z = brief_word<6, DT_BW>(c);
      hi.w = brief_word<7, DT_BW>(c);
      const int f = s_f[k];
      out[2 * (size_t)f] = lo;
      out[2 * (size_t)f + 1] = hi;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(DWARPS * 32) describe_kernel(Geometry g, const uint8_t* __restrict__ blurred,
                                                                const uint32_t* __restrict__ kp_xy,
                                                                const int32_t* __restrict__ n_desc,
                                                                uint8_t* __restrict__ desc, int kp_stride) {
  extern __shared__ __align__(16) uint8_t s_patches[];
  const int img = blockIdx.y;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int n = n_desc[img];
  const int warp = blockIdx.x * DWARPS + wib, n_warps = gridDim.x * DWARPS;
  const uint8_t* base = blurred + (size_t)img * g.rows * g.pitch;
  const uint32_t* xy = kp_xy + (size_t)img * kp_stride;
  uint4* out = reinterpret_cast<uint4*>(desc + (size_t)img * kp_stride * kDescBytes);
  uint8_t* mine = s_patches + (size_t)(wib * 32 + lane) * PSTRIDE;
  uint8_t* warp_patches = s_patches + (size_t)(wib * 32) * PSTRIDE;

  // staging slots of this lane inside one patch: 208 words = 6.5 per lane
  int src_off[7], dst_off[7];
#pragma unroll
  for (int t = 0; t < 7; ++t) {
    const int idx = lane + 32 * t;
    src_off[t] = (idx >> 3) * g.pitch + 4 * (idx & 7);
    dst_off[t] = (idx >> 3) * PP + 4 * (idx & 7);
  }

  for (int i0 = warp * 32; i0 < n; i0 += n_warps * 32) {
    const int i = i0 + lane;
    const uint32_t q = i < n ? xy[i] : 0u;
    const int cnt = min(32, n - i0);
    for (int j = 0; j < cnt; j += 4) {     // four keypoints per step: 28 independent loads in flight per lane
      uint32_t r[4][7];
      const uint8_t* p[4];
      int slot[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        slot[u] = min(j + u, cnt - 1);     // the tail repeats the last keypoint (same data, same destination)
        const uint32_t qq = __shfl_sync(0xffffffffu, q, slot[u]);
        p[u] = base + (size_t)((int)(qq >> 16) - 13) * g.pitch + (((int)(qq & 0xffffu) - 13) & ~3);
      }
#pragma unroll
      for (int t = 0; t < 7; ++t)
        if (t < 6 || lane < 16) {
#pragma unroll
          for (int u = 0; u < 4; ++u) r[u][t] = __ldg(reinterpret_cast<const uint32_t*>(p[u] + src_off[t]));
        }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint8_t* d = warp_patches + (size_t)slot[u] * PSTRIDE;
#pragma unroll
        for (int t = 0; t < 7; ++t)
          if (t < 6 || lane < 16) *reinterpret_cast<uint32_t*>(d + dst_off[t]) = r[u][t];
      }
    }
    __syncwarp();
    if (i < n) {
      const uint8_t* c = mine + (((int)(q & 0xffffu) - 13) & 3);
      uint4 lo, hi;
      lo.x = brief_word<0>(c);
      lo.y = brief_word<1>(c);
      lo.z = brief_word<2>(c);
      lo.w = brief_word<3>(c);
      hi.x = brief_word<4>(c);
      hi.y = brief_word<5>(c);
      hi.z = brief_word<6>(c);
      hi.w = brief_word<7>(c);
      out[2 * (size_t)i] = lo;
      out[2 * (size_t)i + 1] = hi;
    }
    __syncwarp();
  }
}

}  // namespace

bool make_blur_tensor_map(const Geometry& g, const uint8_t* images, int n_images, CUtensorMap* out) {
  return make_image_tensor_map(g, images, n_images, BL_SW, BL_SH, out);
}

// `image_map`: TMA descriptor of b.image (all images of the handle) with the blur tile as box
void launch_blur(const Geometry& g, const Buffers& b, const CUtensorMap& image_map, int first_image, int n_images,
                 cudaStream_t stream) {
  GaussKernel gk;
  {  // cv::getGaussianKernel(7, 2.0, CV_32F)
    double t[7], sum = 0;
    for (int i = 0; i < 7; ++i) {
      const double x = i - 3;
      t[i] = exp(-(x * x) / (2.0 * 2.0 * 2.0));
      sum += t[i];
    }
    for (int i = 0; i < 7; ++i) gk.k[i] = (float)(t[i] / sum);
  }
  const int strips = (g.rows + BL_H - 1) / BL_H;
  dim3 grid((g.cols + BL_W - 1) / BL_W, (strips + BL_WARPS - 1) / BL_WARPS, n_images);
  // (5 or 6 resident CTAs per SM instead of 4 measured the same: the packed-FP pipe, not latency, bounds the walk)
  blur_kernel<<<grid, BL_WARPS * 32, 0, stream>>>(image_map, g, gk, first_image,
                                                   b.blurred + (size_t)first_image * g.rows * g.pitch);
}

// ---- BRIEF-32 (cv::xfeatures2d::BriefDescriptorExtractor::create(32), reference base_framepoint_generator.cpp:186;
// algorithm of opencv_contrib xfeatures2d/src/brief.cpp).  The extractor compares 9 x 9 box sums read from an integral
// image; the same integers come from a box-sum image (<= 81 * 255 fits u16), computed once per frame.

constexpr int BX_W = 128, BX_H = 32;

__global__ void __launch_bounds__(256) box9_kernel(Geometry g, const uint8_t* __restrict__ image,
                                                   uint16_t* __restrict__ boxsum) {
  __shared__ uint8_t s_in[BX_H + 8][BX_W + 16];
  __shared__ uint16_t s_row[BX_H + 8][BX_W];
  const int tid = threadIdx.x, img = blockIdx.z;
  const int x0 = blockIdx.x * BX_W, y0 = blockIdx.y * BX_H;
  const uint8_t* base = image + (size_t)img * g.rows * g.pitch;
  for (int i = tid; i < (BX_H + 8) * (BX_W + 8); i += 256) {
    const int r = i / (BX_W + 8), c = i - r * (BX_W + 8);
    const int gy = y0 - 4 + r, gx = x0 - 4 + c;
    s_in[r][c] = (gy >= 0 && gy < g.rows && gx >= 0 && gx < g.cols) ? base[(size_t)gy * g.pitch + gx] : 0;
  }
  __syncthreads();
  for (int i = tid; i < (BX_H + 8) * BX_W; i += 256) {
    const int r = i / BX_W, c = i - r * BX_W;
    int v = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) v += s_in[r][c + k];
    s_row[r][c] = (uint16_t)v;
  }
  __syncthreads();
  uint16_t* out = boxsum + (size_t)img * g.rows * g.pitch;
  for (int i = tid; i < BX_H * BX_W; i += 256) {
    const int r = i / BX_W, c = i - r * BX_W;
    if (y0 + r >= g.rows || x0 + c >= g.pitch) continue;
    int v = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) v += s_row[r + k][c];
    out[(size_t)(y0 + r) * g.pitch + x0 + c] = (uint16_t)v;
  }
}

// one warp per keypoint, lane b -> descriptor byte b: tests 8b .. 8b+7, bit (7 - k) = test 8b + k (generated_32.i)
__global__ void __launch_bounds__(256) describe_brief_kernel(Geometry g, const uint16_t* __restrict__ boxsum,
                                                             const int8_t* __restrict__ tests,
                                                             const uint32_t* __restrict__ kp_xy,
                                                             const int32_t* __restrict__ n_desc,
                                                             uint8_t* __restrict__ desc, int kp_stride) {
  __shared__ int8_t s_tests[1024];
  const int img = blockIdx.y, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 1024; i += 256) s_tests[i] = tests[i];
  __syncthreads();
  const int n = n_desc[img];
  const uint16_t* S = boxsum + (size_t)img * g.rows * g.pitch;
  const uint32_t* xy = kp_xy + (size_t)img * kp_stride;
  uint8_t* out = desc + (size_t)img * kp_stride * kDescBytes;
  for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += gridDim.x * 8) {
    const uint32_t q = xy[i];
    const int cx = (int)(q & 0xffffu), cy = (int)(q >> 16);   // (int)(pt + 0.5) of integer-valued coordinates
    unsigned v = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int8_t* t = s_tests + (lane * 8 + k) * 4;
      const int a = S[(size_t)(cy + t[0]) * g.pitch + cx + t[1]];
      const int b = S[(size_t)(cy + t[2]) * g.pitch + cx + t[3]];
      v |= (unsigned)(a < b) << (7 - k);
    }
    out[(size_t)i * kDescBytes + lane] = (uint8_t)v;
  }
}

void launch_box9(const Geometry& g, const uint8_t* image, uint16_t* boxsum, int n_images, cudaStream_t stream) {
  dim3 grid((g.cols + BX_W - 1) / BX_W, (g.rows + BX_H - 1) / BX_H, n_images);
  box9_kernel<<<grid, 256, 0, stream>>>(g, image, boxsum);
}

void launch_describe_brief(const Geometry& g, const uint16_t* boxsum, const int8_t* tests, const uint32_t* xy,
                           const int32_t* n, uint8_t* desc, int stride, int n_images, cudaStream_t stream) {
  dim3 grid(n_images <= 8 ? 64 : 16, n_images);
  describe_brief_kernel<<<grid, 256, 0, stream>>>(g, boxsum, tests, xy, n, desc, stride);
}

// opt in to > 48 KB dynamic shared memory once per device
static void configure_describe() {
  static unsigned long long configured = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 64 || !((configured >> dev) & 1ull)) {
    cudaFuncSetAttribute(describe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DSMEM);
    if (dev < 64) configured |= 1ull << dev;
  }
}

// tensor map of an image stack [n_images][rows][pitch] u8 with a box of box_w x box_h x 1 bytes
bool make_image_tensor_map(const Geometry& g, const uint8_t* images, int n_images, int box_w, int box_h, CUtensorMap* out) {
  typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiled encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn ||
        qr != cudaDriverEntryPointSuccess)
      return false;
    encode = reinterpret_cast<EncodeTiled>(fn);
  }
  const cuuint64_t dims[3] = {(cuuint64_t)g.pitch, (cuuint64_t)g.rows, (cuuint64_t)n_images};
  const cuuint64_t strides[2] = {(cuuint64_t)g.pitch, (cuuint64_t)g.pitch * g.rows};   // bytes, dims 1 and 2
  const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t elem[3] = {1, 1, 1};
  return encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(images), dims, strides, box, elem,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// tensor map of a lane's blurred scratch with the describe tile as box
bool make_blurred_tensor_map(const Geometry& g, const uint8_t* blurred, int n_images, CUtensorMap* out) {
  return make_image_tensor_map(g, blurred, n_images, DT_BW, DT_BH, out);
}

// `blurred_map` describes the buffer whose image 0 is image `first_image` of the batch (the lane's scratch)
void launch_describe(const Geometry& g, const Buffers& b, const CUtensorMap& blurred_map, int first_image, int n_images,
                     cudaStream_t stream, int scratch_first) {
  const int tiles_x = (g.cols - 62 + DT_W - 1) / DT_W, tiles_y = (g.rows - 62 + DT_H - 1) / DT_H;
  if (tiles_x <= 0 || tiles_y <= 0) return;   // no pixel is 31 px away from every border: no descriptor-valid keypoint
  dim3 grid(tiles_x, tiles_y, n_images);
  describe_tile_kernel<<<grid, DT_THREADS, 0, stream>>>(blurred_map, g, b.row_ptr + (size_t)first_image * (g.rows + 1),
                                                        b.kp_xy + (size_t)first_image * g.cap,
                                                        b.desc + (size_t)first_image * g.cap * kDescBytes, scratch_first);
}

void launch_describe_at(const Geometry& g, const uint8_t* blurred, const uint32_t* xy, const int32_t* n, uint8_t* desc,
                        int stride, int n_images, cudaStream_t stream) {
  configure_describe();
  dim3 grid(8, n_images);
  describe_kernel<<<grid, DWARPS * 32, DSMEM, stream>>>(g, blurred, xy, n, desc, stride);
}

}  // namespace vslam
