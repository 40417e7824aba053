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
constexpr int TH = 16;
constexpr int HX = 16;
constexpr int SW = TW + 2 * HX;   // 160
constexpr int SH = TH + 6;        // 22

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

__global__ void __launch_bounds__(256) blur_kernel(Geometry g, GaussKernel gk, const uint8_t* __restrict__ image,
                                                   uint8_t* __restrict__ blurred) {
  __shared__ __align__(16) uint8_t s_in[SH][SW];
  __shared__ float s_tmp[SH][TW];
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

  for (int i = tid; i < SH * TW; i += 256) {
    const int r = i / TW, x = i - r * TW;
    const uint8_t* p = &s_in[r][x + HX - 3];
    float acc = __fmul_rn(gk.k[0], (float)p[0]);
#pragma unroll
    for (int t = 1; t < 7; ++t) acc = __fmaf_rn(gk.k[t], (float)p[t], acc);
    s_tmp[r][x] = acc;
  }
  __syncthreads();

  uint8_t* out = blurred + (size_t)img * g.rows * g.pitch;
  for (int i = tid; i < TH * (TW / 4); i += 256) {
    const int y = i / (TW / 4), xq = (i - y * (TW / 4)) * 4;
    if (y0 + y >= g.rows || x0 + xq >= g.pitch) continue;
    uint32_t packed = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int x = xq + j;
      float acc = __fmul_rn(gk.k[3], s_tmp[y + 3][x]);
#pragma unroll
      for (int t = 1; t <= 3; ++t)
        acc = __fmaf_rn(gk.k[3 + t], __fadd_rn(s_tmp[y + 3 + t][x], s_tmp[y + 3 - t][x]), acc);
      int v = __float2int_rn(acc);
      v = min(max(v, 0), 255);
      packed |= (uint32_t)v << (8 * j);
    }
    *reinterpret_cast<uint32_t*>(out + (size_t)(y0 + y) * g.pitch + x0 + xq) = packed;
  }
}

// one warp per keypoint, lane b produces descriptor byte b (tests 8b..8b+7, LSB first)
__global__ void __launch_bounds__(256) describe_kernel(Geometry g, const uint8_t* __restrict__ blurred,
                                                       const uint32_t* __restrict__ kp_xy,
                                                       const int32_t* __restrict__ n_desc, uint8_t* __restrict__ desc) {
  const int img = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int n_warps = gridDim.x * 8;
  const int n = n_desc[img];
  if (warp >= n) return;
  int o0[8], o1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int8_t* p = &c_pattern[(lane * 8 + k) * 4];
    o0[k] = p[1] * g.pitch + p[0];
    o1[k] = p[3] * g.pitch + p[2];
  }
  const uint8_t* base = blurred + (size_t)img * g.rows * g.pitch;
  const uint32_t* xy = kp_xy + (size_t)img * g.cap;
  uint8_t* out = desc + (size_t)img * g.cap * kDescBytes;
  for (int i = warp; i < n; i += n_warps) {
    const uint32_t q = xy[i];
    const uint8_t* c = base + (size_t)(q >> 16) * g.pitch + (q & 0xffffu);
    unsigned v = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) v |= (unsigned)(c[o0[k]] < c[o1[k]]) << k;
    out[(size_t)i * kDescBytes + lane] = (uint8_t)v;
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
