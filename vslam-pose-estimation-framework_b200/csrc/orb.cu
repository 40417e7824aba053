// orb.cu -- K3 blur_kernel (7x7 sigma 2 Gaussian, float separable, rounded to u8) + K4 describe_kernel (rBRIEF-256).
//
// Replaces cv::ORB::create()->compute(image, keypoints, descriptors) as called by
// BaseFramePointGenerator::computeDescriptors (reference src/framepoint_generation/base_framepoint_generator.cpp:
// 195, 222, 431-438).  Semantics: SURVEY.md Appendix A.3.  The float evaluation order is the contract shared with
// the CPU oracle (oracle/c/vslam_oracle.c, orc_gauss7_u8):
//   row pass   : acc = k[0]*p[-3]; acc = fma(k[i], p[i-3], acc), i = 1..6
//   column pass: acc = k[3]*r[0];  acc = fma(k[3+j], r[+j] + r[-j], acc), j = 1..3 ; out = rint(acc)
#include "kernels.cuh"

namespace vslam {

namespace {

constexpr int TW = 128;
constexpr int TH = 32;
constexpr int HX = 16;
constexpr int SW = TW + 2 * HX;   // 160
constexpr int SH = TH + 6;        // 38

__constant__ int8_t c_pattern[256 * 4] = {
#include "orb_pattern_31.inc"
};

struct GaussKernel {
  float k[7];
};

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
  return i;
}

// exact u8 -> float without the (quarter-rate) I2F unit: place the byte in the mantissa of 2^23 and subtract 2^23
__device__ __forceinline__ float byte_to_float(uint32_t word, int j) {
  return __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7650u | (unsigned)j)) - 8388608.0f;
}

// K3.  Tile 128 x 32 outputs.  Row pass: one thread per 4 adjacent outputs (10 input bytes converted once, 28 FMA);
// column pass: one thread per 4 columns x 4 rows (10 float4 shared-memory loads, packed 32-bit stores).
__global__ void __launch_bounds__(256) blur_kernel(Geometry g, GaussKernel gk, const uint8_t* __restrict__ image,
                                                   uint8_t* __restrict__ blurred) {
  __shared__ __align__(16) uint8_t s_in[SH][SW];
  __shared__ __align__(16) float s_tmp[SH][TW];
  const int tid = threadIdx.x;
  const int img = blockIdx.z;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const uint8_t* base = image + (size_t)img * g.rows * g.pitch;

  for (int i = tid; i < SH * (SW / 16); i += 256) {
    const int r = i / (SW / 16), c = i - r * (SW / 16);
    const int gy = reflect101(y0 - 3 + r, g.rows);
    const int gx = x0 - HX + c * 16;
    const uint8_t* row = base + (size_t)gy * g.pitch;
    if (gx >= 0 && gx + 16 <= g.cols) {
      *reinterpret_cast<uint4*>(&s_in[r][c * 16]) = __ldg(reinterpret_cast<const uint4*>(row + gx));
    } else {
      for (int j = 0; j < 16; ++j) {
        const int x = gx + j;
        // only columns within the 3 px filter support of this tile are ever read
        s_in[r][c * 16 + j] = (x >= x0 - 3 && x < x0 + TW + 3) ? row[reflect101(x, g.cols)] : 0;
      }
    }
  }
  __syncthreads();

  // ---- row pass: outputs x = xq .. xq+3 need inputs xq-3 .. xq+6, all inside three aligned words
  for (int i = tid; i < SH * (TW / 4); i += 256) {
    const int r = i / (TW / 4), xq = (i - r * (TW / 4)) * 4;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&s_in[r][xq + HX - 4]);
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
    float p[10];
    p[0] = byte_to_float(w0, 1);
    p[1] = byte_to_float(w0, 2);
    p[2] = byte_to_float(w0, 3);
#pragma unroll
    for (int j = 0; j < 4; ++j) p[3 + j] = byte_to_float(w1, j);
    p[7] = byte_to_float(w2, 0);
    p[8] = byte_to_float(w2, 1);
    p[9] = byte_to_float(w2, 2);
    float4 out;
    float* o = reinterpret_cast<float*>(&out);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float acc = __fmul_rn(gk.k[0], p[j]);
#pragma unroll
      for (int t = 1; t < 7; ++t) acc = __fmaf_rn(gk.k[t], p[j + t], acc);
      o[j] = acc;
    }
    *reinterpret_cast<float4*>(&s_tmp[r][xq]) = out;
  }
  __syncthreads();

  // ---- column pass + round half-to-even to u8 (adding 1.5 * 2^23 leaves rint(acc) in the low mantissa bits)
  uint8_t* outp = blurred + (size_t)img * g.rows * g.pitch;
  {
    const int xq = (tid & 31) * 4, yb = (tid >> 5) * 4;
    float4 t[10];
#pragma unroll
    for (int r = 0; r < 10; ++r) t[r] = *reinterpret_cast<const float4*>(&s_tmp[yb + r][xq]);
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      if (y0 + yb + y >= g.rows || x0 + xq >= g.pitch) continue;
      uint32_t packed = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float* c = reinterpret_cast<const float*>(&t[y + 3]) + j;
        float acc = __fmul_rn(gk.k[3], *c);
#pragma unroll
        for (int d = 1; d <= 3; ++d) {
          const float up = reinterpret_cast<const float*>(&t[y + 3 + d])[j];
          const float dn = reinterpret_cast<const float*>(&t[y + 3 - d])[j];
          acc = __fmaf_rn(gk.k[3 + d], __fadd_rn(up, dn), acc);
        }
        acc = fminf(fmaxf(acc, 0.0f), 255.0f);
        packed |= (__float_as_uint(__fadd_rn(acc, 12582912.0f)) & 0xffu) << (8 * j);
      }
      *reinterpret_cast<uint32_t*>(outp + (size_t)(y0 + yb + y) * g.pitch + x0 + xq) = packed;
    }
  }
}

// K4.  One warp per keypoint, lane b produces descriptor byte b (tests 8b..8b+7, LSB first).
// The 26 x 26 patch the 512 test points fall into (offsets -13..12) is staged in shared memory with coalesced
// 32-bit loads (8 aligned words per row cover x-13..x+12 at any alignment) and the 16 byte look-ups per lane then
// hit shared memory instead of scattering over ~26 L1 sectors each; the next keypoint's patch is prefetched into
// registers while the current one is evaluated.
constexpr int PR = 26;       // patch rows / columns
constexpr int PP = 36;       // patch row pitch in shared memory (bytes): 9 words, spreads rows over banks
constexpr int PW = 8;        // aligned words loaded per patch row
constexpr int PL = (PR * PW + 31) / 32;   // 7 loads per lane

__global__ void __launch_bounds__(256) describe_kernel(Geometry g, const uint8_t* __restrict__ blurred,
                                                       const uint32_t* __restrict__ kp_xy,
                                                       const int32_t* __restrict__ n_desc, uint8_t* __restrict__ desc) {
  __shared__ __align__(16) uint8_t s_patch[8][PR * PP];
  const int img = blockIdx.y;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = blockIdx.x * 8 + wib;
  const int n_warps = gridDim.x * 8;
  const int n = n_desc[img];
  if (warp >= n) return;
  int o0[8], o1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int8_t* p = &c_pattern[(lane * 8 + k) * 4];
    o0[k] = (p[1] + 13) * PP + p[0] + 13;
    o1[k] = (p[3] + 13) * PP + p[2] + 13;
  }
  int src_off[PL], dst_off[PL];   // per-lane word slots of the patch
#pragma unroll
  for (int t = 0; t < PL; ++t) {
    const int idx = lane + 32 * t;
    const int r = idx / PW, w = idx - r * PW;
    src_off[t] = idx < PR * PW ? r * g.pitch + 4 * w : -1;
    dst_off[t] = r * PP + 4 * w;
  }
  const uint8_t* base = blurred + (size_t)img * g.rows * g.pitch;
  const uint32_t* xy = kp_xy + (size_t)img * g.cap;
  uint8_t* out = desc + (size_t)img * g.cap * kDescBytes;
  uint8_t* patch = s_patch[wib];

  uint32_t regs[PL];
  auto fetch = [&](uint32_t q) {
    const int x = (int)(q & 0xffffu) - 13, y = (int)(q >> 16) - 13;
    const uint8_t* p = base + (size_t)y * g.pitch + (x & ~3);
#pragma unroll
    for (int t = 0; t < PL; ++t)
      if (src_off[t] >= 0) regs[t] = __ldg(reinterpret_cast<const uint32_t*>(p + src_off[t]));
  };
  uint32_t q = xy[warp];
  fetch(q);
  for (int i = warp; i < n; i += n_warps) {
#pragma unroll
    for (int t = 0; t < PL; ++t)
      if (src_off[t] >= 0) *reinterpret_cast<uint32_t*>(patch + dst_off[t]) = regs[t];
    __syncwarp();
    const int shift = ((int)(q & 0xffffu) - 13) & 3;
    const int nxt = i + n_warps;
    if (nxt < n) {
      q = xy[nxt];
      fetch(q);
    }
    const uint8_t* c = patch + shift;
    unsigned v = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) v |= (unsigned)(c[o0[k]] < c[o1[k]]) << k;
    out[(size_t)i * kDescBytes + lane] = (uint8_t)v;
    __syncwarp();
  }
}

}  // namespace

void launch_blur(const Geometry& g, const Buffers& b, int first_image, int n_images, cudaStream_t stream) {
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
  dim3 grid((g.cols + TW - 1) / TW, (g.rows + TH - 1) / TH, n_images);
  blur_kernel<<<grid, 256, 0, stream>>>(g, gk, b.image + (size_t)first_image * g.rows * g.pitch,
                                        b.blurred + (size_t)first_image * g.rows * g.pitch);
}

void launch_describe(const Geometry& g, const Buffers& b, int first_image, int n_images, cudaStream_t stream) {
  dim3 grid(n_images <= 8 ? 32 : 8, n_images);
  describe_kernel<<<grid, 256, 0, stream>>>(g, b.blurred + (size_t)first_image * g.rows * g.pitch,
                                            b.kp_xy + (size_t)first_image * g.cap, b.n_desc + first_image,
                                            b.desc + (size_t)first_image * g.cap * kDescBytes);
}

}  // namespace vslam
