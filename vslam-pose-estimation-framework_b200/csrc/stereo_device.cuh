// stereo_device.cuh -- device helpers shared by the stereo matching (stereo.cu) and the tracking / recovery kernels
// (track.cu).  Both translation units are compiled with -fmad=false: the expression order below is the parity
// contract with the CPU oracle.
#pragma once

#include "kernels.cuh"

namespace vslam {

// stereo_framepoint_generator.cpp:109-125 ; 0.1 * SRRG_PROSLAM_DESCRIPTOR_SIZE_BITS with 256 bits
__device__ __forceinline__ double triangulation_threshold(const StereoParams& sp, int n_left) {
  const double tenth = __dmul_rn(0.1, 256.0);
  if (sp.localizing) return fmin(tenth, sp.max_matching_distance);
  const double ratio = fmin(__ddiv_rn((double)n_left, (double)sp.target_keypoints), 1.0);
  return fmax(__dmul_rn(ratio, sp.max_matching_distance), tenth);
}

__device__ __forceinline__ int popc256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
         __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// stereo_framepoint_generator.cpp:871-895 ; x, y are integer-valued floats
__device__ __forceinline__ void triangulate(const StereoParams& sp, float xl, float yl, float xr, float yr,
                                            double out[3]) {
  const double z = __ddiv_rn(sp.bx, (double)__fsub_rn(xr, xl));
  out[0] = __dmul_rn(__dmul_rn(__ddiv_rn(1.0, sp.fx), __dsub_rn((double)xl, sp.cx)), z);
  out[1] = __dmul_rn(__dmul_rn(__ddiv_rn(1.0, sp.fy), __dsub_rn(__ddiv_rn((double)__fadd_rn(yl, yr), 2.0), sp.cy)), z);
  out[2] = z;
}

// `const int32_t v = <double expression>;` as the reference's x86-64 build evaluates it (cvttsd2si): truncation toward
// zero, NaN and out-of-range values give INT32_MIN (which every caller rejects as `< 0`)
__device__ __forceinline__ int32_t to_i32(double v) {
  if (!(v > -2147483649.0 && v < 2147483648.0)) return INT32_MIN;
  return __double2int_rz(v);
}

}  // namespace vslam
