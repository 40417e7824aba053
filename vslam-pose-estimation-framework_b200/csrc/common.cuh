// common.cuh -- shared device-side types of libvslam_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace vslam {

constexpr int kMaxRegions = 64;
constexpr int kDescBytes = 32;

// one FAST detector region (cv::Rect, base_framepoint_generator.cpp:293-296) + its integer threshold
struct Region {
  int x, y, w, h;
};

// geometry shared by every kernel of the framepoint path; passed by value (kernel parameter space)
struct Geometry {
  int rows, cols;
  int pitch;        // bytes between image rows on the device (multiple of 128)
  int mask_words;   // uint32 words per keypoint-mask row (multiple of 4)
  int cap;          // descriptor-valid keypoints per image
  int n_regions;
  int rows_bin, cols_bin, bin_size;
  int enable_binning;
  int border;       // keypoints closer than this to the image border get no descriptor: 31 (cv::ORB) / 28 (BRIEF-32)
};

// device-resident state of a batch of stereo pairs; image index = 2*pair + side
struct Buffers {
  uint8_t* image;        // [2B][rows][pitch]
  uint8_t* blurred;      // [2B][rows][pitch]      ORB: 7x7 Gaussian (u8) | BRIEF-32: 9x9 box sums (u16, 2 bytes per pixel)
  uint32_t* mask;        // [2B][rows][mask_words]   raw FAST keypoints (after NMS)
  int32_t* raw_count;    // [2B][n_regions]
  int32_t* row_ptr;      // [2B][rows+1]             CSR over rows of the descriptor-valid keypoints
  uint32_t* kp_xy;       // [2B][cap]                x | y << 16, sorted by (row, col)
  uint8_t* kp_score;     // [2B][cap]
  uint8_t* desc;         // [2B][cap][32]            sorted order
  int32_t* n_desc;       // [2B]
  int2* match;           // [B][cap]   per sorted left feature: x = sorted right index or -1, y = dist | pass << 16
  uint8_t* pruned_l;     // [B][cap]   left features consumed by tracking or matched in an earlier epipolar pass
  uint8_t* consumed_r;   // [B][cap]   same for the right features
  int32_t* n_out;        // [B][2]     {n_framepoints, n_matches}
  int32_t* error_flag;   // [1]        != 0: capacity exceeded
};

}  // namespace vslam
