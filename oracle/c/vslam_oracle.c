/*
 * vslam_oracle.c -- CPU restatement ("Tier A" oracle) of the reference hot path.  See vslam_oracle.h for
 * the status header (TEST INFRASTRUCTURE ONLY; pinned against oracle/_ref, the reference's own translation units).
 *
 * Compile with -ffp-contract=off: every fused multiply-add below is an explicit fmaf()/none, so that the
 * float blur and the double triangulation have ONE defined evaluation order (the CUDA kernels use the
 * same order with explicit __fmaf_rn / __dmul_rn / __ddiv_rn).
 *
 * All file:line citations are relative to /root/reference.
 */
#include "vslam_oracle.h"

#include <stdio.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * detector regions, bins, thresholds
 * ---------------------------------------------------------------------------------------------- */

/* src/framepoint_generation/base_framepoint_generator.cpp:231-299 */
void orc_detector_regions(int rows, int cols, int nv, int nh, orc_rect* out) {
  const double pixel_rows_per_detector = (double)rows / nv;  /* :234 */
  const double pixel_cols_per_detector = (double)cols / nh;  /* :235 */
  for (int r = 0; r < nv; ++r) {
    for (int c = 0; c < nh; ++c) {
      int offset_width = 0, offset_height = 0;               /* :263-270 */
      if (nv > 1) offset_height = 2;
      if (nh > 1) offset_width = 2;
      int offset_r = 0, offset_c = 0;                        /* :273-290 */
      if (r > 0) {
        offset_r = -offset_height;
        if (r < nv - 1) offset_height *= 2;
      }
      if (c > 0) {
        offset_c = -offset_width;
        if (c < nh - 1) offset_width *= 2;
      }
      /* :293-296  cv::Rect(int,int,int,int) from doubles -> truncation */
      orc_rect q;
      q.x = (int)(round(c * pixel_cols_per_detector) + offset_c);
      q.y = (int)(round(r * pixel_rows_per_detector) + offset_r);
      q.w = (int)(pixel_cols_per_detector + offset_width);
      q.h = (int)(pixel_rows_per_detector + offset_height);
      out[r * nh + c] = q;
    }
  }
}

/* base_framepoint_generator.cpp:304-305 */
void orc_bin_grid(int rows, int cols, int bin_size, int* rows_bin, int* cols_bin) {
  *cols_bin = (int)(floor((double)cols / bin_size) + 1);
  *rows_bin = (int)(floor((double)rows / bin_size) + 1);
}

/* base_framepoint_generator.cpp:377-415 */
double orc_threshold_proposal(double threshold, int n_keypoints, double target, double tolerance,
                              double max_change, double thr_min, double thr_max) {
  double detector_threshold = threshold;
  const double delta = ((double)n_keypoints - target) / target;              /* :382 */
  if (delta < -tolerance) {                                                   /* :385 */
    const double change = fmax(delta, -max_change);                           /* :388 */
    detector_threshold = detector_threshold + fmin(change * detector_threshold, -1.0); /* :391 */
    if (detector_threshold < thr_min) detector_threshold = thr_min;           /* :394 */
  } else if (delta > tolerance) {                                             /* :400 */
    const double change = fmin(delta, max_change);                            /* :403 */
    detector_threshold += fmax(change * detector_threshold, 1.0);             /* :406 */
    if (detector_threshold > thr_max) detector_threshold = thr_max;           /* :409 */
  }
  return detector_threshold;
}

/* base_framepoint_generator.cpp:440-459 with _number_of_detections == 2 (left + right image) */
void orc_adjust_thresholds(double* thresholds, int n_regions, const int* counts_l, const int* counts_r,
                           double target, double tolerance, double max_change, double thr_min, double thr_max) {
  for (int i = 0; i < n_regions; ++i) {
    double acc = 0;
    acc += orc_threshold_proposal(thresholds[i], counts_l[i], target, tolerance, max_change, thr_min, thr_max);
    acc += orc_threshold_proposal(thresholds[i], counts_r[i], target, tolerance, max_change, thr_min, thr_max);
    acc /= 2;                      /* :449 */
    thresholds[i] = rint(acc);     /* :451 -> FastDetector::setThreshold -> std::rint (:21) */
  }
}

/* ------------------------------------------------------------------------------------------------
 * FAST-9/16 with score and 3x3 non-maximum suppression == cv::FastFeatureDetector::create(t)->detect
 * (OpenCV is not vendored by the reference; call sites base_framepoint_generator.cpp:10-24,367;
 *  algorithm restated from SURVEY.md Appendix A.1, pinned against cv2 4.13 in the tests)
 * ---------------------------------------------------------------------------------------------- */

static const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

static int fast_score(const uint8_t* p, int stride, int t) {
  int d[25];
  const int v = p[0];
  {
    /* quick reject: every 9-arc contains ring pixel 0 or 8 (and 4 or 12) -- pure speed-up, same result */
    const int d0 = v - p[3 * stride], d8 = v - p[-3 * stride], d4 = v - p[3], d12 = v - p[-3];
    if (abs(d0) <= t && abs(d8) <= t) return 0;
    if (abs(d4) <= t && abs(d12) <= t) return 0;
  }
  for (int k = 0; k < 16; ++k) d[k] = v - p[kRingDy[k] * stride + kRingDx[k]];
  for (int k = 0; k < 9; ++k) d[16 + k] = d[k];
  int A = -1000, Bm = 1000;
  for (int s = 0; s < 16; ++s) {
    int mn = d[s], mx = d[s];
    for (int k = 1; k < 9; ++k) {
      if (d[s + k] < mn) mn = d[s + k];
      if (d[s + k] > mx) mx = d[s + k];
    }
    if (mn > A) A = mn;
    if (mx < Bm) Bm = mx;
  }
  if (!(A > t || Bm < -t)) return 0;
  int a0 = t > A ? t : A;          /* cornerScore<16>: a0 = max(threshold, max-min over dark arcs) */
  int b0 = -a0 < Bm ? -a0 : Bm;    /* b0 = min(-a0, min-max over bright arcs) */
  return -b0 - 1;
}

int orc_fast_detect(const uint8_t* img, int stride, int w, int h, int threshold, orc_kp* out, int cap) {
  if (threshold < 0) threshold = 0;
  if (threshold > 255) threshold = 255;
  if (w < 7 || h < 7) return 0;
  int* score = (int*)calloc((size_t)w * h, sizeof(int));
  for (int y = 3; y <= h - 4; ++y)
    for (int x = 3; x <= w - 4; ++x) score[y * w + x] = fast_score(img + (size_t)y * stride + x, stride, threshold);
  int n = 0;
  for (int y = 3; y <= h - 4; ++y) {
    for (int x = 3; x <= w - 4; ++x) {
      const int s = score[y * w + x];
      if (s == 0) continue;
      const int* q = score + y * w + x;
      if (s > q[-1] && s > q[1] && s > q[-w - 1] && s > q[-w] && s > q[-w + 1] && s > q[w - 1] && s > q[w] &&
          s > q[w + 1]) {
        if (n < cap) {
          out[n].x = (float)x;
          out[n].y = (float)y;
          out[n].response = (float)s;
        }
        ++n;
      }
    }
  }
  free(score);
  return n;
}

/* base_framepoint_generator.cpp:362-424 : per-region detection, shift by region.tl(), concatenate */
int orc_detect_keypoints(const uint8_t* img, int stride, int rows, int cols, int nv, int nh,
                         const double* thresholds, orc_kp* out, int cap, int* counts) {
  orc_rect* regions = (orc_rect*)malloc(sizeof(orc_rect) * nv * nh);
  orc_detector_regions(rows, cols, nv, nh, regions);
  int n = 0;
  for (int i = 0; i < nv * nh; ++i) {
    const orc_rect q = regions[i];
    const int room = cap - n > 0 ? cap - n : 0;
    const int m = orc_fast_detect(img + (size_t)q.y * stride + q.x, stride, q.w, q.h, (int)rint(thresholds[i]),
                                  out + (n < cap ? n : cap), room);
    const int wrote = m < room ? m : room;
    for (int k = 0; k < wrote; ++k) {          /* :418-419 */
      out[n + k].x += (float)q.x;
      out[n + k].y += (float)q.y;
    }
    if (counts) counts[i] = m;
    n += m;
  }
  free(regions);
  return n;
}

/* ------------------------------------------------------------------------------------------------
 * ORB descriptor (cv::ORB::create()->compute on FAST keypoints), SURVEY.md Appendix A.3
 * ---------------------------------------------------------------------------------------------- */

static const int8_t kOrbPattern[256 * 4] = {
#include "../data/orb_pattern_31.inc"
};

/* cv::getGaussianKernel(7, 2.0, CV_32F) */
void orc_gauss7_kernel(float k[7]) {
  double t[7], sum = 0;
  for (int i = 0; i < 7; ++i) {
    const double x = i - 3;
    t[i] = exp(-(x * x) / (2.0 * 2.0 * 2.0));
    sum += t[i];
  }
  for (int i = 0; i < 7; ++i) k[i] = (float)(t[i] / sum);
}

static inline int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) {
    if (i < 0) i = -i;
    else i = 2 * n - 2 - i;
  }
  return i;
}

/* Float separable 7-tap Gaussian, BORDER_REFLECT_101, rounded half-to-even to u8.
 * Defined evaluation order (the contract the CUDA kernel reproduces bit-for-bit):
 *   row pass   : acc = k[0]*p[-3]; acc = fmaf(k[i], p[i-3], acc) for i = 1..6           (left to right)
 *   column pass: acc = k[3]*r[0];  acc = fmaf(k[3+j], r[+j] + r[-j], acc) for j = 1..3  (symmetric)
 *   out = (uint8) rintf(acc) saturated                                                                   */
void orc_gauss7_u8(const uint8_t* img, int stride, int w, int h, uint8_t* out, int out_stride) {
  float k[7];
  orc_gauss7_kernel(k);
  float* tmp = (float*)malloc(sizeof(float) * (size_t)w * h);
  /* the interior runs without the border look-up (same taps, same order: results identical, loops vectorisable) */
  for (int y = 0; y < h; ++y) {
    const uint8_t* row = img + (size_t)y * stride;
    float* t = tmp + (size_t)y * w;
    for (int x = 0; x < w; ++x) {
      if (x == 3 && w > 6) {
        for (; x < w - 3; ++x) {
          float acc = k[0] * (float)row[x - 3];
          for (int i = 1; i < 7; ++i) acc = fmaf(k[i], (float)row[x + i - 3], acc);
          t[x] = acc;
        }
      }
      float acc = k[0] * (float)row[reflect101(x - 3, w)];
      for (int i = 1; i < 7; ++i) acc = fmaf(k[i], (float)row[reflect101(x + i - 3, w)], acc);
      t[x] = acc;
    }
  }
  for (int y = 0; y < h; ++y) {
    const float* c = tmp + (size_t)y * w;
    const float* u[3];
    const float* d[3];
    for (int j = 1; j <= 3; ++j) {
      d[j - 1] = tmp + (size_t)reflect101(y + j, h) * w;
      u[j - 1] = tmp + (size_t)reflect101(y - j, h) * w;
    }
    uint8_t* o = out + (size_t)y * out_stride;
    for (int x = 0; x < w; ++x) {
      float acc = k[3] * c[x];
      for (int j = 1; j <= 3; ++j) {
        const float s = d[j - 1][x] + u[j - 1][x];
        acc = fmaf(k[3 + j], s, acc);
      }
      float r = rintf(acc);
      if (r < 0) r = 0;
      if (r > 255) r = 255;
      o[x] = (uint8_t)r;
    }
  }
  free(tmp);
}

int orc_orb_compute(const uint8_t* img, int stride, int w, int h, const uint8_t* blurred, int bstride,
                    orc_kp* kps, int n_kps, uint8_t* desc) {
  uint8_t* own = NULL;
  if (!blurred) {
    own = (uint8_t*)malloc((size_t)w * h);
    orc_gauss7_u8(img, stride, w, h, own, w);
    blurred = own;
    bstride = w;
  }
  int n = 0;
  for (int i = 0; i < n_kps; ++i) {
    /* KeyPointsFilter::runByImageBorder(edgeThreshold = 31): keep [31, w-31) x [31, h-31) */
    const float x = kps[i].x, y = kps[i].y;
    if (!(x >= 31 && x < w - 31 && y >= 31 && y < h - 31)) continue;
    kps[n] = kps[i];
    const uint8_t* c = blurred + (size_t)(int)lrintf(y) * bstride + (int)lrintf(x);
    uint8_t* d = desc + (size_t)n * 32;
    for (int b = 0; b < 32; ++b) {
      unsigned v = 0;
      for (int k = 0; k < 8; ++k) {
        const int8_t* p = kOrbPattern + (b * 8 + k) * 4;
        const int t0 = c[p[1] * bstride + p[0]];
        const int t1 = c[p[3] * bstride + p[2]];
        v |= (unsigned)(t0 < t1) << k;
      }
      d[b] = (uint8_t)v;
    }
    ++n;
  }
  free(own);
  return n;
}

/* xfeatures2d/src/brief.cpp (opencv_contrib): smoothedSum() + pixelTests32() + BriefDescriptorExtractorImpl::compute() */
int orc_brief32_compute(const uint8_t* img, int stride, int w, int h, const int8_t* tests, orc_kp* kps, int n_kps,
                        uint8_t* desc) {
  /* cv::integral(image, sum, CV_32S): sum is (h+1) x (w+1), sum(y, x) = sum of img[0..y) x [0..x) */
  int32_t* sum = (int32_t*)calloc((size_t)(h + 1) * (w + 1), sizeof(int32_t));
  const int sp = w + 1;
  for (int y = 0; y < h; ++y) {
    int32_t run = 0;
    for (int x = 0; x < w; ++x) {
      run += img[(size_t)y * stride + x];
      sum[(size_t)(y + 1) * sp + x + 1] = sum[(size_t)y * sp + x + 1] + run;
    }
  }
  const int HALF_KERNEL = 4, border = 48 / 2 + 9 / 2;   /* PATCH_SIZE = 48, KERNEL_SIZE = 9 */
  int n = 0;
  for (int i = 0; i < n_kps; ++i) {
    const float x = kps[i].x, y = kps[i].y;
    if (!(x >= border && x < w - border && y >= border && y < h - border)) continue;   /* runByImageBorder */
    kps[n] = kps[i];
    const int cy = (int)(y + 0.5f), cx = (int)(x + 0.5f);
    uint8_t* d = desc + (size_t)n * 32;
    for (int b = 0; b < 32; ++b) {
      unsigned v = 0;
      for (int k = 0; k < 8; ++k) {
        const int8_t* t = tests + (b * 8 + k) * 4;
        int s[2];
        for (int j = 0; j < 2; ++j) {
          const int iy = cy + t[2 * j], ix = cx + t[2 * j + 1];
          s[j] = sum[(size_t)(iy + HALF_KERNEL + 1) * sp + ix + HALF_KERNEL + 1] -
                 sum[(size_t)(iy + HALF_KERNEL + 1) * sp + ix - HALF_KERNEL] -
                 sum[(size_t)(iy - HALF_KERNEL) * sp + ix + HALF_KERNEL + 1] +
                 sum[(size_t)(iy - HALF_KERNEL) * sp + ix - HALF_KERNEL];
        }
        v |= (unsigned)(s[0] < s[1]) << (7 - k);
      }
      d[b] = (uint8_t)v;
    }
    ++n;
  }
  free(sum);
  return n;
}

int orc_hamming256(const uint8_t* a, const uint8_t* b) {
  int d = 0;
  for (int i = 0; i < 32; ++i) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
  return d;
}

/* ------------------------------------------------------------------------------------------------
 * stereo matching, triangulation, binning
 * ---------------------------------------------------------------------------------------------- */

/* stereo_framepoint_generator.cpp:109-125 (SRRG_PROSLAM_DESCRIPTOR_SIZE_BITS = 256, CMakeLists.txt:32) */
double orc_triangulation_threshold(int localizing, int n_left, int target, double max_dist) {
  if (localizing) return fmin(0.1 * 256, max_dist);
  const double ratio = fmin((double)n_left / target, 1.0);
  return fmax(ratio * max_dist, 0.1 * 256);
}

static int feature_less(const void* a, const void* b) {
  const orc_feature* fa = (const orc_feature*)a;
  const orc_feature* fb = (const orc_feature*)b;
  if (fa->row != fb->row) return fa->row < fb->row ? -1 : 1;
  if (fa->col != fb->col) return fa->col < fb->col ? -1 : 1;
  return fa->index < fb->index ? -1 : (fa->index > fb->index);
}

/* intensity_feature_matcher.cpp:48-70 (setFeatures) + :72-79 (sortFeatureVector); frame_point.h:24-30 */
void orc_make_features(const orc_kp* kps, const uint8_t* desc, int n, orc_feature* out) {
  for (int i = 0; i < n; ++i) {
    out[i].x = kps[i].x;
    out[i].y = kps[i].y;
    out[i].row = (int32_t)kps[i].y;
    out[i].col = (int32_t)kps[i].x;
    out[i].index = i;
    memcpy(out[i].desc, desc + (size_t)i * 32, 32);
  }
  qsort(out, (size_t)n, sizeof(orc_feature), feature_less);
}

/* stereo_framepoint_generator.cpp:871-895 ; pt coordinates are float (cv::Point2f) */
void orc_triangulate(const orc_stereo_camera* cam, float xl, float yl, float xr, float yr, double out[3]) {
  const double z = cam->bx / (double)(xr - xl);                                /* :882-883, float subtraction */
  const double x = ((1 / cam->fx) * ((double)xl - cam->cx)) * z;               /* :886-887 */
  const double y = ((1 / cam->fy) * ((double)(yl + yr) / 2.0 - cam->cy)) * z;  /* :890-893, float addition */
  out[0] = x;
  out[1] = y;
  out[2] = z;
}

typedef struct {
  int32_t match;        /* >= 0: index into matches; < 0: -(k+1) tracked point k; INT32_MIN: empty */
  int has_previous;
  double disparity, distance;
} bin_cell;

static int prune(orc_feature* f, int n, const uint8_t* matched) {  /* intensity_feature_matcher.cpp:150-172 */
  int m = 0;
  for (int i = 0; i < n; ++i)
    if (!matched[i]) f[m++] = f[i];
  return m;
}

int orc_stereo_compute(orc_feature* fl, int* n_l, orc_feature* fr, int* n_r, const orc_stereo_camera* cam,
                       double max_distance, double min_disparity, int max_epipolar_offset, int enable_binning,
                       int bin_size, int rows, int cols, const orc_tracked* tracked, int n_tracked,
                       orc_match* matches, int32_t* winners, int* n_winners) {
  int rows_bin = 0, cols_bin = 0;
  bin_cell* bins = NULL;
  if (enable_binning) {
    orc_bin_grid(rows, cols, bin_size, &rows_bin, &cols_bin);
    bins = (bin_cell*)malloc(sizeof(bin_cell) * (size_t)rows_bin * cols_bin);
    for (int i = 0; i < rows_bin * cols_bin; ++i) bins[i].match = INT32_MIN;
    for (int k = 0; k < n_tracked; ++k) {                                      /* :147-155 */
      const int rb = (int)rint((double)tracked[k].row / bin_size);
      const int cb = (int)rint((double)tracked[k].col / bin_size);
      bin_cell* c = &bins[rb * cols_bin + cb];
      c->match = -(k + 1);
      c->has_previous = tracked[k].has_previous;
      c->disparity = tracked[k].disparity;
      c->distance = tracked[k].distance;
    }
  }

  int number_of_new_points = 0;
  int nl = *n_l, nr = *n_r;
  uint8_t* matched_l = (uint8_t*)malloc((size_t)(nl > 0 ? nl : 1));
  uint8_t* matched_r = (uint8_t*)malloc((size_t)(nr > 0 ? nr : 1));

  const int n_offsets = 1 + 2 * max_epipolar_offset;                           /* :45-50 */
  for (int oi = 0; oi < n_offsets; ++oi) {
    const int epipolar_offset = oi == 0 ? 0 : ((oi & 1) ? (oi + 1) / 2 : -(oi / 2));
    memset(matched_l, 0, (size_t)(nl > 0 ? nl : 1));
    memset(matched_r, 0, (size_t)(nr > 0 ? nr : 1));
    int index_R = 0;                                                           /* :289 */
    for (int index_L = 0; index_L < nl; index_L++) {                           /* :292 */
      if (index_R == nr) break;                                                /* :294 */
      while (fl[index_L].row < fr[index_R].row + epipolar_offset) {            /* :299-305 */
        index_L++;
        if (index_L == nl) break;
      }
      if (index_L == nl) break;                                                /* :306 */
      const orc_feature* feature_left = &fl[index_L];
      while (feature_left->row > fr[index_R].row + epipolar_offset) {          /* :312-318 */
        index_R++;
        if (index_R == nr) break;
      }
      if (index_R == nr) break;                                                /* :319 */

      int index_search_R = index_R;                                            /* :324-327 */
      double descriptor_distance_best = max_distance;
      int index_best_R = 0;
      while (feature_left->row == fr[index_search_R].row + epipolar_offset) {  /* :330 */
        if (feature_left->col - fr[index_search_R].col < 0) break;             /* :333 */
        const double descriptor_distance = orc_hamming256(feature_left->desc, fr[index_search_R].desc);
        if (descriptor_distance < descriptor_distance_best) {                  /* :342 */
          descriptor_distance_best = descriptor_distance;
          index_best_R = index_search_R;
        }
        index_search_R++;
        if (index_search_R == nr) break;                                       /* :347 */
      }

      if (descriptor_distance_best < max_distance) {                           /* :353 */
        const orc_feature* feature_right = &fr[index_best_R];
        if (feature_left->col - feature_right->col < min_disparity) continue;  /* :358-361 */

        orc_match* m = &matches[number_of_new_points];
        m->index_left = feature_left->index;
        m->index_right = feature_right->index;
        m->xl = feature_left->x;
        m->yl = feature_left->y;
        m->xr = feature_right->x;
        m->yr = feature_right->y;
        m->distance = (int32_t)descriptor_distance_best;
        m->epipolar_offset = epipolar_offset;
        orc_triangulate(cam, m->xl, m->yl, m->xr, m->yr, m->cam);

        if (enable_binning) {                                                  /* :371-394 */
          const int rb = (int)rint((double)feature_left->row / bin_size);
          const int cb = (int)rint((double)feature_left->col / bin_size);
          bin_cell* c = &bins[rb * cols_bin + cb];
          const double disparity = (double)(m->xl - m->xr);                    /* frame_point.cpp:19 (float sub) */
          if (c->match != INT32_MIN) {
            if (!c->has_previous && disparity > c->disparity && descriptor_distance_best <= c->distance) {
              c->match = number_of_new_points;
              c->disparity = disparity;
              c->distance = descriptor_distance_best;
            }
          } else {
            c->match = number_of_new_points;
            c->has_previous = 0;
            c->disparity = disparity;
            c->distance = descriptor_distance_best;
          }
        }
        ++number_of_new_points;                                                /* :397-398 */
        matched_l[index_L] = 1;                                                /* :405-406 */
        matched_r[index_best_R] = 1;
        index_R = index_best_R + 1;                                            /* :414 */
      }
    }
    nl = prune(fl, nl, matched_l);                                             /* :419-420 */
    nr = prune(fr, nr, matched_r);
  }
  *n_l = nl;
  *n_r = nr;

  int nw = 0;
  if (enable_binning) {                                                        /* :435-455 */
    for (int rb = 0; rb < rows_bin; ++rb)
      for (int cb = 0; cb < cols_bin; ++cb) {
        const bin_cell* c = &bins[rb * cols_bin + cb];
        if (c->match != INT32_MIN && !c->has_previous) winners[nw++] = c->match;
      }
  } else {                                                                     /* :456-460 */
    for (int i = 0; i < number_of_new_points; ++i) winners[nw++] = i;
  }
  *n_winners = nw;
  free(bins);
  free(matched_l);
  free(matched_r);
  return number_of_new_points;
}

/* ------------------------------------------------------------------------------------------------
 * tracking (track) and recovery (recoverPoints)
 * ---------------------------------------------------------------------------------------------- */

/* `const int32_t v = <double expression>;` as x86-64 evaluates it (cvttsd2si): truncation toward zero; NaN and
 * values outside int32 give INT32_MIN ("integer indefinite"), which every caller below rejects as `< 0`. */
static int32_t to_i32(double v) {
  if (!(v > -2147483649.0 && v < 2147483648.0)) return INT32_MIN;
  return (int32_t)v;
}

/* intensity_feature_matcher.cpp:81-148 on a lattice of feature positions (-1 = nullptr).  Returns the position
 * of the chosen feature or -1; *distance_best as the reference leaves descriptor_distance_best_. */
static int match_in_region(const int32_t* lattice, int cols, const orc_feature* f, int row_reference,
                           int col_reference, const uint8_t* descriptor_reference, int row_start, int row_end,
                           int col_start, int col_end, double maximum_distance, int track_by_appearance,
                           double* distance_best) {
  *distance_best = maximum_distance;                                           /* :91 */
  int best = -1;
  if (track_by_appearance) {                                                   /* :96-112 */
    for (int row = row_start; row < row_end; ++row)
      for (int col = col_start; col < col_end; ++col) {
        const int32_t k = lattice[(size_t)row * cols + col];
        if (k < 0) continue;
        const double d = orc_hamming256(descriptor_reference, f[k].desc);
        if (d < *distance_best) {
          *distance_best = d;
          best = k;
        }
      }
  } else {                                                                     /* :115-139 */
    uint32_t projection_distance_best = 10000;
    for (int row = row_start; row < row_end; ++row)
      for (int col = col_start; col < col_end; ++col) {
        const int32_t k = lattice[(size_t)row * cols + col];
        if (k < 0) continue;
        const double d = orc_hamming256(descriptor_reference, f[k].desc);
        if (d < maximum_distance) {
          const int32_t dr = row_reference - row, dc = col_reference - col;
          const uint32_t projection_distance = (uint32_t)(dr * dr + dc * dc);
          if (projection_distance < projection_distance_best) {
            projection_distance_best = projection_distance;
            *distance_best = d;
            best = k;
          }
        }
      }
  }
  return best;                                                                 /* :142-147 */
}

static int32_t* make_lattice(const orc_feature* f, int n, int rows, int cols) {  /* :48-70 setFeatures */
  int32_t* lattice = (int32_t*)malloc(sizeof(int32_t) * (size_t)rows * cols);
  for (size_t i = 0; i < (size_t)rows * cols; ++i) lattice[i] = -1;
  for (int i = 0; i < n; ++i) lattice[(size_t)f[i].row * cols + f[i].col] = i;
  return lattice;
}

static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

/* stereo_framepoint_generator.cpp:464-681.  Floating-point evaluation order of the two Eigen products is not pinned
 * by the reference (Eigen is un-vendored): T*p is evaluated row by row, left to right, like transform_point below;
 * K*p as fx*X + cx*Z, fy*Y + cy*Z, Z (the zero products of the dense 3x3 product add exact zeros). */
int orc_track(const orc_feature* fl, int n_l, const orc_feature* fr, int n_r, int rows, int cols,
              const orc_stereo_camera* cam, const orc_previous_point* previous, int n_previous, const double T[12],
              int track_by_appearance, int tracking_distance_pixels, double max_distance_tracking,
              double max_distance_triangulation, double min_disparity, orc_track_record* tracks, int32_t* lost,
              int* n_lost, uint8_t* matched_l, uint8_t* matched_r, int* n_tracked_landmarks,
              double* accumulated_distance) {
  int32_t* lattice_l = make_lattice(fl, n_l, rows, cols);
  int32_t* lattice_r = make_lattice(fr, n_r, rows, cols);
  memset(matched_l, 0, (size_t)(n_l > 0 ? n_l : 1));
  memset(matched_r, 0, (size_t)(n_r > 0 ? n_r : 1));
  int number_of_tracked_points = 0, number_of_points_lost = 0, tracked_landmarks = 0;
  double accumulated = 0;
  const int D = tracking_distance_pixels;

  for (int u = 0; u < n_previous; ++u) {                                       /* :494 */
    const orc_previous_point* pp = &previous[u];
    double pc[3];
    for (int i = 0; i < 3; ++i)                                                /* :496-498 */
      pc[i] = T[4 * i] * pp->cam[0] + T[4 * i + 1] * pp->cam[1] + T[4 * i + 2] * pp->cam[2] + T[4 * i + 3];
    const double il[3] = {cam->fx * pc[0] + cam->cx * pc[2], cam->fy * pc[1] + cam->cy * pc[2], pc[2]}; /* :501-502 */
    const int32_t col_l = to_i32(il[0] / il[2]);                               /* :503-506 */
    const int32_t row_l = to_i32(il[1] / il[2]);
    if (col_l < 0 || col_l > cols || row_l < 0 || row_l > rows) continue;      /* :509-514 */

    double distance_best = max_distance_tracking;                              /* :517 */
    int row_start = imax(row_l - D, 0), row_end = imin(row_l + D + 1, rows);   /* :520-529 */
    int col_start = imax(col_l - D, 0), col_end = imin(col_l + D + 1, cols);
    const int kl = match_in_region(lattice_l, cols, fl, row_l, col_l, pp->desc_left, row_start, row_end, col_start,
                                   col_end, max_distance_tracking, track_by_appearance, &distance_best); /* :532-538 */
    int tracked = 0;
    if (kl >= 0) {
      const orc_feature* feature_left = &fl[kl];
      const float error_x = (float)col_l - feature_left->x;                    /* :543-545 cv::Point2f */
      const float error_y = (float)row_l - feature_left->y;
      const double ir[3] = {il[0] + cam->bx, il[1] + 0.0, il[2] + 0.0};        /* :549 _baseline = (b_x, 0, 0) */
      const int32_t col_r = to_i32(ir[0] / ir[2] - (double)error_x);           /* :550-555 */
      const int32_t row_r = to_i32(ir[1] / ir[2] - (double)error_y);
      if (col_r < 0 || col_r > cols || row_r < 0 || row_r > rows) continue;    /* :558-563 */
      const int32_t e = (int32_t)fabs((double)pp->epipolar_offset);            /* :568-569 */
      row_start = imax(row_r - e, 0);                                          /* :570-579 */
      row_end = imin(row_r + e + 1, rows);
      col_start = imax(col_r - D, 0);
      col_end = imin(col_r + D + 1, feature_left->col);
      const int kr = match_in_region(lattice_r, cols, fr, row_r, col_r, feature_left->desc, row_start, row_end,
                                     col_start, col_end, max_distance_triangulation, 1, &distance_best); /* :584-590 */
      if (kr >= 0) {
        const orc_feature* feature_right = &fr[kr];
        if ((double)(feature_left->col - feature_right->col) < min_disparity) continue;                  /* :597-600 */
        if ((double)orc_hamming256(feature_right->desc, pp->desc_right) > max_distance_tracking) continue; /* :603-607 */
        for (int col = feature_right->col + 1; col < feature_left->col; ++col) {                         /* :611-620 */
          int32_t* cell = &lattice_r[(size_t)feature_right->row * cols + col];
          if (*cell >= 0) {
            matched_r[*cell] = 1;
            *cell = -1;
          }
        }
        orc_track_record* t = &tracks[number_of_tracked_points++];                    /* :623-643 */
        memset(t, 0, sizeof(*t));
        t->index_previous = u;
        t->index_left = feature_left->index;
        t->index_right = feature_right->index;
        t->xl = feature_left->x; t->yl = feature_left->y;
        t->xr = feature_right->x; t->yr = feature_right->y;
        t->distance = (int32_t)distance_best;
        t->epipolar_offset = feature_right->row - feature_left->row;
        t->projection_left[0] = (float)col_l; t->projection_left[1] = (float)row_l;
        t->projection_right[0] = (float)(ir[0] / ir[2]); t->projection_right[1] = (float)(ir[1] / ir[2]);
        t->projection_right_corrected[0] = (float)col_r; t->projection_right_corrected[1] = (float)row_r;
        orc_triangulate(cam, t->xl, t->yl, t->xr, t->yr, t->cam);
        accumulated += distance_best;                                          /* :627 */
        matched_l[kl] = 1;                                                     /* :646-651 */
        matched_r[kr] = 1;
        lattice_l[(size_t)feature_left->row * cols + feature_left->col] = -1;
        lattice_r[(size_t)feature_right->row * cols + feature_right->col] = -1;
        if (pp->has_landmark) ++tracked_landmarks;                             /* :653-655 */
        tracked = 1;
      }
    }
    if (!tracked) lost[number_of_points_lost++] = u;                           /* :660-663 : !point_previous->next() */
  }
  free(lattice_l);
  free(lattice_r);
  *n_lost = number_of_points_lost;
  *n_tracked_landmarks = tracked_landmarks;
  *accumulated_distance = accumulated;   /* :666-667 divides by number_of_tracked_points (NaN when zero) */
  return number_of_tracked_points;
}

/* rBRIEF-256 at an integer pixel of a blurred image (the inner loop of orc_orb_compute) */
static void brief_at(const uint8_t* blurred, int stride, int x, int y, uint8_t* d) {
  const uint8_t* c = blurred + (size_t)y * stride + x;
  for (int b = 0; b < 32; ++b) {
    unsigned v = 0;
    for (int k = 0; k < 8; ++k) {
      const int8_t* p = kOrbPattern + (b * 8 + k) * 4;
      v |= (unsigned)((int)c[p[1] * stride + p[0]] < (int)c[p[3] * stride + p[2]]) << k;
    }
    d[b] = (uint8_t)v;
  }
}

/* BRIEF-32 at an integer pixel, box sums taken directly from the image (== the integral-image differences) */
static void brief32_at(const uint8_t* img, int stride, int x, int y, const int8_t* tests, uint8_t* d) {
  for (int b = 0; b < 32; ++b) {
    unsigned v = 0;
    for (int k = 0; k < 8; ++k) {
      const int8_t* t = tests + (b * 8 + k) * 4;
      int s[2] = {0, 0};
      for (int j = 0; j < 2; ++j)
        for (int dy = -4; dy <= 4; ++dy)
          for (int dx = -4; dx <= 4; ++dx) s[j] += img[(size_t)(y + t[2 * j] + dy) * stride + x + t[2 * j + 1] + dx];
      v |= (unsigned)(s[0] < s[1]) << (7 - k);
    }
    d[b] = (uint8_t)v;
  }
}

/* stereo_framepoint_generator.cpp:683-869.  brief_tests == NULL: ORB extractor, blurred_l / blurred_r are the blurred
 * frames; otherwise BRIEF-32 with that test table and blurred_l / blurred_r are the RAW frames. */
int orc_recover_points(const uint8_t* blurred_l, const uint8_t* blurred_r, int stride, int rows, int cols,
                       const orc_stereo_camera* cam, const orc_previous_point* lost, int n_lost,
                       const double W[12], double min_depth, double max_depth, double max_distance_tracking,
                       double max_distance_triangulation, double min_disparity, const int8_t* brief_tests,
                       orc_recovered* out) {
  int n = 0;
  for (int u = 0; u < n_lost; ++u) {                                           /* :702 */
    const orc_previous_point* pp = &lost[u];
    if (!pp->has_landmark) continue;                                           /* :704-706 */
    double pc[3];
    for (int i = 0; i < 3; ++i)                                                /* :716-717 */
      pc[i] = W[4 * i] * pp->world[0] + W[4 * i + 1] * pp->world[1] + W[4 * i + 2] * pp->world[2] + W[4 * i + 3];
    double il[3] = {cam->fx * pc[0] + cam->cx * pc[2], cam->fy * pc[1] + cam->cy * pc[2], pc[2]};        /* :725-726 */
    double ir[3] = {il[0] + cam->bx, il[1] + 0.0, il[2] + 0.0};                                          /* :727-728 */
    if (il[2] < min_depth || il[2] > max_depth || ir[2] < min_depth || ir[2] > max_depth) continue;     /* :731-736 */
    /* :739-740  `v /= v.z()` : Eigen evaluates the scalar first; element-wise division */
    const double zl = il[2], zr = ir[2];
    for (int i = 0; i < 3; ++i) { il[i] /= zl; ir[i] /= zr; }
    const float plx = (float)rint(il[0]), ply = (float)rint(il[1]);            /* :743-746 */
    const float prx = (float)rint(ir[0]), pry = (float)rint(ir[1]);
    const float border = 5 * pp->keypoint_size;                                /* :749-750 */
    if (plx < border + 1 || plx > cols - border - 1 || prx < border + 1 || prx > cols - border - 1 ||
        ply < border + 1 || ply > rows - border - 1 || pry < border + 1 || pry > rows - border - 1) continue; /* :751-766 */
    /* :769-795 / :808-822 : cv::ORB::compute on the (2*border+1)^2 region around the projection with the keypoint at
     * its centre; with border >= 31 the keypoint survives ORB's 31 px border filter and its patch + blur support
     * are interior, so the descriptor is rBRIEF of the blurred frame at the projection */
    if (border < (brief_tests ? 28 : 31)) continue;                            /* descriptor.rows == 0 (:790-792) */
    orc_recovered r;
    memset(&r, 0, sizeof(r));
    if (brief_tests) brief32_at(blurred_l, stride, (int)plx, (int)ply, brief_tests, r.desc_left);
    else brief_at(blurred_l, stride, (int)plx, (int)ply, r.desc_left);
    if ((double)orc_hamming256(pp->desc_left, r.desc_left) > max_distance_tracking) continue;            /* :798-802 */
    if (brief_tests) brief32_at(blurred_r, stride, (int)prx, (int)pry, brief_tests, r.desc_right);
    else brief_at(blurred_r, stride, (int)prx, (int)pry, r.desc_right);
    if ((double)(plx - prx) < min_disparity) continue;                                                   /* :825-828 */
    if ((double)orc_hamming256(pp->desc_right, r.desc_right) > max_distance_tracking) continue;          /* :831-835 */
    const int d = orc_hamming256(r.desc_left, r.desc_right);                                             /* :838-843 */
    if ((double)d > max_distance_triangulation) continue;
    r.index_lost = u;
    r.distance = d;
    r.xl = plx; r.yl = ply; r.xr = prx; r.yr = pry;
    orc_triangulate(cam, plx, ply, prx, pry, r.cam);                                                     /* :851-855 */
    out[n++] = r;
  }
  return n;
}

/* ------------------------------------------------------------------------------------------------
 * aligners
 * ---------------------------------------------------------------------------------------------- */

static void transform_point(const double T[12], const double* p, double out[3]) {
  for (int i = 0; i < 3; ++i) out[i] = T[4 * i] * p[0] + T[4 * i + 1] * p[1] + T[4 * i + 2] * p[2] + T[4 * i + 3];
}

/* jt = [wt*I3 | -2*skew(p)] ; skew(p) = [[0,-z,y],[z,0,-x],[-y,x,0]] (srrg_core, un-vendored) */
static void jacobian_transform(const double p[3], double wt, double jt[18]) {
  memset(jt, 0, sizeof(double) * 18);
  jt[0 * 6 + 0] = wt;
  jt[1 * 6 + 1] = wt;
  jt[2 * 6 + 2] = wt;
  jt[0 * 6 + 4] = -2 * -p[2];
  jt[0 * 6 + 5] = -2 * p[1];
  jt[1 * 6 + 3] = -2 * p[2];
  jt[1 * 6 + 5] = -2 * -p[0];
  jt[2 * 6 + 3] = -2 * -p[1];
  jt[2 * 6 + 4] = -2 * p[0];
}

static void mat3x3_times_3x6(const double K[9], const double jt[18], double out[18]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 6; ++j) out[i * 6 + j] = K[i * 3] * jt[j] + K[i * 3 + 1] * jt[6 + j] + K[i * 3 + 2] * jt[12 + j];
}

static void sys_reset(orc_linear_system* s) { memset(s, 0, sizeof(*s)); }

/* H += J^T W J ; b += J^T W e, with W = diag(w[0..D)) */
static void accumulate(orc_linear_system* s, const double* J, const double* w, const double* e, int D) {
  for (int i = 0; i < 6; ++i) {
    for (int j = 0; j < 6; ++j) {
      double a = 0;
      for (int d = 0; d < D; ++d) a += J[d * 6 + i] * w[d] * J[d * 6 + j];
      s->H[i * 6 + j] += a;
    }
    double a = 0;
    for (int d = 0; d < D; ++d) a += J[d * 6 + i] * w[d] * e[d];
    s->b[i] += a;
  }
}

/* src/aligners/stereouv_aligner.cpp:72-187 */
void orc_stereouv_linearize(const orc_aligner_problem* p, const double T[12], int ignore_outliers,
                            orc_linear_system* sys, double* errors, uint8_t* inliers) {
  sys_reset(sys);                                                              /* :75-78 */
  for (int u = 0; u < p->n; ++u) {
    errors[u] = -1;                                                            /* :82-84 */
    inliers[u] = 0;
    double omega = p->omega[u];
    double pc[3];
    transform_point(T, p->moving + 3 * u, pc);                                 /* :87 */
    if (pc[2] < p->min_depth) continue;                                        /* :88 */
    double abc_l[3], abc_r[3];
    for (int i = 0; i < 3; ++i) {                                              /* :93-94 */
      abc_l[i] = p->K[3 * i] * pc[0] + p->K[3 * i + 1] * pc[1] + p->K[3 * i + 2] * pc[2];
      abc_r[i] = abc_l[i] + p->baseline[i];
    }
    const double c_l = abc_l[2], c_r = abc_r[2];
    const double ul = abc_l[0] / c_l, vl = abc_l[1] / c_l;                     /* :99-100 */
    const double ur = abc_r[0] / c_r, vr = abc_r[1] / c_r;
    if (ul < 0 || ul > p->cols || vl < 0 || vl > p->rows) continue;            /* :103-106 */
    if (ur < 0 || ur > p->cols || vr < 0 || vr > p->rows) continue;            /* :107-110 */
    const double* f = p->fixed + 4 * u;
    const double e[4] = {ul - f[0], vl - f[1], ur - f[2], vr - f[3]};          /* :115-118 */
    const double chi = omega * e[0] * e[0] + omega * e[1] * e[1] + omega * e[2] * e[2] + omega * e[3] * e[3]; /* :121 */
    errors[u] = chi;                                                           /* :124 */
    if (chi > p->kernel) {                                                     /* :127-133 */
      if (ignore_outliers) continue;
      omega *= p->kernel / chi;
    } else {
      inliers[u] = 1;
      ++sys->inliers;
    }
    sys->total_error += errors[u];                                             /* :140 */
    double jt[18], kj[18], J[24];
    jacobian_transform(pc, p->wt[u], jt);                                      /* :143-149 */
    mat3x3_times_3x6(p->K, jt, kj);                                            /* :152 */
    const double il = 1 / c_l, ir = 1 / c_r;                                   /* :155-158 */
    const double il2 = il * il, ir2 = ir * ir;
    for (int j = 0; j < 6; ++j) {                                              /* :161-177 */
      J[0 * 6 + j] = il * kj[j] + (-abc_l[0] * il2) * kj[12 + j];
      J[1 * 6 + j] = il * kj[6 + j] + (-abc_l[1] * il2) * kj[12 + j];
      J[2 * 6 + j] = ir * kj[j] + (-abc_r[0] * ir2) * kj[12 + j];
      J[3 * 6 + j] = ir * kj[6 + j] + (-abc_r[1] * ir2) * kj[12 + j];
    }
    const double w[4] = {omega, omega, omega, omega};
    accumulate(sys, J, w, e, 4);                                               /* :183-184 */
  }
  sys->outliers = p->n - sys->inliers;                                         /* :186 */
}

/* src/aligners/uvd_aligner.cpp:77-171 */
void orc_uvd_linearize(const orc_aligner_problem* p, const double T[12], int ignore_outliers,
                       orc_linear_system* sys, double* errors, uint8_t* inliers) {
  sys_reset(sys);
  for (int u = 0; u < p->n; ++u) {
    errors[u] = -1;                                                            /* :88-90 */
    inliers[u] = 0;
    double w[3] = {p->omega[2 * u], p->omega[2 * u], p->omega[2 * u + 1]};
    double pc[3];
    transform_point(T, p->moving + 3 * u, pc);                                 /* :93 */
    const double depth = pc[2];
    if (depth <= p->min_depth) continue;                                       /* :95 */
    double uvd[3];
    for (int i = 0; i < 3; ++i) uvd[i] = p->K[3 * i] * pc[0] + p->K[3 * i + 1] * pc[1] + p->K[3 * i + 2] * pc[2]; /* :100 */
    const double px = uvd[0] / uvd[2], py = uvd[1] / uvd[2];                   /* :103 */
    if (px < 0 || px > p->cols || py < 0 || py > p->rows) continue;            /* :109-112 */
    const double* f = p->fixed + 3 * u;
    const double e[3] = {px - f[0], py - f[1], depth - f[2]};                  /* :115-117 */
    const double chi = w[0] * e[0] * e[0] + w[1] * e[1] * e[1] + w[2] * e[2] * e[2]; /* :120 */
    errors[u] = chi;
    if (chi > p->kernel) {                                                     /* :126-135 */
      if (ignore_outliers) continue;
      const double s = p->kernel / chi;
      w[0] *= s;
      w[1] *= s;
      w[2] *= s;
    } else {
      inliers[u] = 1;
      ++sys->inliers;
    }
    sys->total_error += errors[u];                                             /* :138 */
    const double iz = 1 / depth, iz2 = iz * iz;                                /* :141-142 */
    double jt[18], kj[18], J[18];
    jacobian_transform(pc, p->wt[u], jt);                                      /* :145-152 */
    mat3x3_times_3x6(p->K, jt, kj);                                            /* :161 (jacobian_projection*K*jt) */
    for (int j = 0; j < 6; ++j) {                                              /* :155-161 */
      J[0 * 6 + j] = iz * kj[j] + (-uvd[0] * iz2) * kj[12 + j];
      J[1 * 6 + j] = iz * kj[6 + j] + (-uvd[1] * iz2) * kj[12 + j];
      J[2 * 6 + j] = kj[12 + j];
    }
    accumulate(sys, J, w, e, 3);                                               /* :167-168 */
  }
  sys->outliers = p->n - sys->inliers;                                         /* :170 */
}

/* Eigen::FullPivLU<Matrix6>::solve restated: LU with complete pivoting (Eigen 3.3.4, un-vendored) */
void orc_solve6_fullpiv(const double A_in[36], const double rhs[6], double x[6]) {
  double A[36], b[6];
  int colperm[6];
  memcpy(A, A_in, sizeof(A));
  memcpy(b, rhs, sizeof(b));
  for (int i = 0; i < 6; ++i) colperm[i] = i;
  int rank = 6;
  double maxpivot = 0;
  for (int k = 0; k < 6; ++k) {
    int pr = k, pc = k;
    double best = -1;
    for (int i = k; i < 6; ++i)
      for (int j = k; j < 6; ++j)
        if (fabs(A[i * 6 + j]) > best) {
          best = fabs(A[i * 6 + j]);
          pr = i;
          pc = j;
        }
    if (best == 0) {
      rank = k;
      break;
    }
    if (best > maxpivot) maxpivot = best;
    if (best > maxpivot) maxpivot = best;
    if (pr != k) {
      for (int j = 0; j < 6; ++j) {
        const double t = A[k * 6 + j];
        A[k * 6 + j] = A[pr * 6 + j];
        A[pr * 6 + j] = t;
      }
      const double t = b[k];
      b[k] = b[pr];
      b[pr] = t;
    }
    if (pc != k) {
      for (int i = 0; i < 6; ++i) {
        const double t = A[i * 6 + k];
        A[i * 6 + k] = A[i * 6 + pc];
        A[i * 6 + pc] = t;
      }
      const int t = colperm[k];
      colperm[k] = colperm[pc];
      colperm[pc] = t;
    }
    for (int i = k + 1; i < 6; ++i) {
      const double f = A[i * 6 + k] / A[k * 6 + k];
      A[i * 6 + k] = f;
      for (int j = k + 1; j < 6; ++j) A[i * 6 + j] -= f * A[k * 6 + j];
      b[i] -= f * b[k];
    }
  }
  {  /* Eigen::FullPivLU::rank(): only pivots above |largest pivot| * epsilon * size are used by solve() */
    int r = 0;
    for (int i = 0; i < rank; ++i) r += fabs(A[i * 6 + i]) > maxpivot * (2.220446049250313e-16 * 6);
    rank = r;
  }
  double y[6] = {0, 0, 0, 0, 0, 0};
  for (int i = rank - 1; i >= 0; --i) {
    double s = b[i];
    for (int j = i + 1; j < rank; ++j) s -= A[i * 6 + j] * y[j];
    y[i] = s / A[i * 6 + i];
  }
  for (int i = 0; i < 6; ++i) x[colperm[i]] = y[i];
}

/* srrg_core::v2t (un-vendored, branch marchless; restated from SURVEY.md section 8c / Appendix B):
 * translation v[0:3]; rotation from the quaternion (w = sqrt(1-|q|^2), q = v[3:6]) if |q|^2 < 1, else (0, q/|q|). */
void orc_v2t(const double v[6], double T[12]) {
  double qx = v[3], qy = v[4], qz = v[5], qw;
  const double n2 = qx * qx + qy * qy + qz * qz;
  if (n2 < 1) {
    qw = sqrt(1 - n2);
  } else {
    const double n = sqrt(n2);
    qx /= n;
    qy /= n;
    qz /= n;
    qw = 0;
  }
  /* Eigen::Quaternion::toRotationMatrix */
  const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
  const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
  const double txx = tx * qx, txy = ty * qx, txz = tz * qx;
  const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  T[0] = 1 - (tyy + tzz); T[1] = txy - twz;       T[2] = txz + twy;       T[3] = v[0];
  T[4] = txy + twz;       T[5] = 1 - (txx + tzz); T[6] = tyz - twx;       T[7] = v[1];
  T[8] = txz - twy;       T[9] = tyz + twx;       T[10] = 1 - (txx + tyy); T[11] = v[2];
}

static void compose(const double A[12], const double B[12], double C[12]) {  /* C = A * B (isometries) */
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) C[4 * i + j] = A[4 * i] * B[j] + A[4 * i + 1] * B[4 + j] + A[4 * i + 2] * B[8 + j];
    C[4 * i + 3] = A[4 * i] * B[3] + A[4 * i + 1] * B[7] + A[4 * i + 2] * B[11] + A[4 * i + 3];
  }
}

/* stereouv_aligner.cpp:190-207 / uvd_aligner.cpp:174-191 */
void orc_one_round(int kind, const orc_aligner_problem* p, double damping, double T[12], int ignore_outliers,
                   orc_linear_system* sys, double* errors, uint8_t* inliers) {
  if (kind == 0) orc_stereouv_linearize(p, T, ignore_outliers, sys, errors, inliers);   /* :193 */
  else orc_uvd_linearize(p, T, ignore_outliers, sys, errors, inliers);
  for (int i = 0; i < 6; ++i) sys->H[i * 6 + i] += damping * p->n;                      /* :196 */
  double nb[6], dx[6], D[12], Tn[12];
  for (int i = 0; i < 6; ++i) nb[i] = -sys->b[i];
  orc_solve6_fullpiv(sys->H, nb, dx);                                                   /* :199 */
  orc_v2t(dx, D);
  compose(D, T, Tn);                                                                    /* :200 */
  /* :203-206  R -= 0.5 * R * (R^T R - I) */
  double RtR[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      RtR[3 * i + j] = Tn[i] * Tn[j] + Tn[4 + i] * Tn[4 + j] + Tn[8 + i] * Tn[8 + j];
      if (i == j) RtR[3 * i + j] -= 1;
    }
  memcpy(T, Tn, sizeof(double) * 12);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      T[4 * i + j] = Tn[4 * i + j] - 0.5 * (Tn[4 * i] * RtR[j] + Tn[4 * i + 1] * RtR[3 + j] + Tn[4 * i + 2] * RtR[6 + j]);
}

/* stereouv_aligner.cpp:210-264 / uvd_aligner.cpp:194-248 */
int orc_converge(int kind, const orc_aligner_problem* p, double damping, double error_delta, int max_iterations,
                 int min_inliers, double T[12], orc_linear_system* sys, double* errors, uint8_t* inliers,
                 double* info, int* rounds) {
  double total_error_previous = 0;
  int converged = 0, n_rounds = 0;
  /* UVD hard-codes 100 (uvd_aligner.cpp:208); StereoUV uses minimum_number_of_inliers (stereouv_aligner.cpp:224) */
  const int inlier_gate = kind == 0 ? min_inliers : 100;
  for (int it = 0; it < max_iterations; ++it) {
    orc_one_round(kind, p, damping, T, 0, sys, errors, inliers);
    ++n_rounds;
    if (error_delta > fabs(total_error_previous - sys->total_error)) {
      total_error_previous = sys->total_error;
      if (sys->inliers > inlier_gate && sys->inliers > sys->outliers) {
        for (int ii = 0; ii < max_iterations; ++ii) {
          orc_one_round(kind, p, damping, T, 1, sys, errors, inliers);
          ++n_rounds;
          if (fabs(total_error_previous - sys->total_error) < error_delta) {
            total_error_previous = sys->total_error;
            break;
          } else {
            total_error_previous = sys->total_error;
          }
        }
      }
      if (info) memcpy(info, sys->H, sizeof(double) * 36);
      converged = 1;
      break;
    } else {
      total_error_previous = sys->total_error;
    }
  }
  if (rounds) *rounds = n_rounds;
  return converged;
}

/* ================================================================================================================
 * Landmark::update (src/types/landmark.cpp:66-152) and the trajectory wire formats
 * (src/types/world_map.cpp:183-252).  Eigen is un-vendored: the evaluation order of its fixed-size products is not
 * pinned by the reference; the order below (left to right, no contraction) is the contract shared with the kernel.
 * ================================================================================================================ */
void orc_solve3_fullpiv(const double A_in[9], const double rhs[3], double x[3]) {
  double A[9], b[3];
  int colperm[3] = {0, 1, 2};
  memcpy(A, A_in, sizeof(A));
  memcpy(b, rhs, sizeof(b));
  int rank = 3;
  double maxpivot = 0;
  for (int k = 0; k < 3; ++k) {
    int pr = k, pc = k;
    double best = -1;
    for (int i = k; i < 3; ++i)
      for (int j = k; j < 3; ++j)
        if (fabs(A[i * 3 + j]) > best) {
          best = fabs(A[i * 3 + j]);
          pr = i;
          pc = j;
        }
    if (best == 0) {
      rank = k;
      break;
    }
    if (best > maxpivot) maxpivot = best;
    if (best > maxpivot) maxpivot = best;
    if (pr != k) {
      for (int j = 0; j < 3; ++j) {
        const double t = A[k * 3 + j];
        A[k * 3 + j] = A[pr * 3 + j];
        A[pr * 3 + j] = t;
      }
      const double t = b[k];
      b[k] = b[pr];
      b[pr] = t;
    }
    if (pc != k) {
      for (int i = 0; i < 3; ++i) {
        const double t = A[i * 3 + k];
        A[i * 3 + k] = A[i * 3 + pc];
        A[i * 3 + pc] = t;
      }
      const int t = colperm[k];
      colperm[k] = colperm[pc];
      colperm[pc] = t;
    }
    for (int i = k + 1; i < 3; ++i) {
      const double f = A[i * 3 + k] / A[k * 3 + k];
      A[i * 3 + k] = f;
      for (int j = k + 1; j < 3; ++j) A[i * 3 + j] -= f * A[k * 3 + j];
      b[i] -= f * b[k];
    }
  }
  {  /* Eigen::FullPivLU::rank(): only pivots above |largest pivot| * epsilon * size are used by solve() */
    int r = 0;
    for (int i = 0; i < rank; ++i) r += fabs(A[i * 3 + i]) > maxpivot * (2.220446049250313e-16 * 3);
    rank = r;
  }
  double y[3] = {0, 0, 0};
  for (int i = rank - 1; i >= 0; --i) {
    double s = b[i];
    for (int j = i + 1; j < rank; ++j) s -= A[i * 3 + j] * y[j];
    y[i] = s / A[i * 3 + i];
  }
  for (int i = 0; i < 3; ++i) x[colperm[i]] = y[i];
}

int orc_landmark_update(const orc_landmark_measurement* ms, int n, const double* w2c, const double* c2w,
                        uint32_t max_iterations, double max_err2, double world[3], uint32_t* number_of_updates,
                        int* iterations) {
  double x[3] = {world[0], world[1], world[2]};                                  /* :82 */
  double total_previous = 0;                                                     /* :88 */
  int outcome = 0;
  uint32_t it = 0;
  for (; it < max_iterations; ++it) {                                            /* :91 */
    double H[9] = {0}, b[3] = {0}, total = 0;                                    /* :92-94 */
    uint32_t outliers = 0;
    for (int m = 0; m < n; ++m) {                                                /* :98 */
      const double* W = w2c + 12 * (size_t)ms[m].frame;
      double p[3], e[3];
      for (int r = 0; r < 3; ++r)                                                /* :102 worldToCameraLeft * x */
        p[r] = ((W[4 * r] * x[0] + W[4 * r + 1] * x[1]) + W[4 * r + 2] * x[2]) + W[4 * r + 3];
      if (p[2] <= 0) {                                                           /* :103-106 */
        ++outliers;
        continue;
      }
      for (int r = 0; r < 3; ++r) e[r] = p[r] - ms[m].camera_coordinates[r];      /* :109 */
      double w = ms[m].inverse_depth_meters;                                     /* :112 omega = I * inverse depth */
      const double err2 = ((e[0] * w) * e[0] + (e[1] * w) * e[1]) + (e[2] * w) * e[2];   /* :115 */
      total += err2;                                                             /* :116 */
      if (err2 > max_err2) {                                                     /* :119-122 */
        w *= max_err2 / err2;
        ++outliers;
      }
      /* :125-132  J = R;  H += J^T (w I) J;  b += J^T (w I) e */
      for (int i = 0; i < 3; ++i) {
        const double jw0 = W[i] * w, jw1 = W[4 + i] * w, jw2 = W[8 + i] * w;       /* row i of J^T * omega */
        for (int j = 0; j < 3; ++j) H[3 * i + j] += (jw0 * W[j] + jw1 * W[4 + j]) + jw2 * W[8 + j];
        b[i] += (jw0 * e[0] + jw1 * e[1]) + jw2 * e[2];
      }
    }
    double nb[3] = {-b[0], -b[1], -b[2]}, dx[3];
    orc_solve3_fullpiv(H, nb, dx);                                               /* :136 */
    for (int i = 0; i < 3; ++i) x[i] += dx[i];
    if (fabs(total - total_previous) < 1e-5 || it == 999) {                      /* :139 */
      const uint32_t inliers = (uint32_t)n - outliers;                           /* :140 (unsigned, as the reference) */
      outcome = 3;
      if (inliers > *number_of_updates) {                                        /* :143-147 */
        world[0] = x[0];
        world[1] = x[1];
        world[2] = x[2];
        *number_of_updates = inliers;
        outcome = 1;
      } else if (inliers < outliers) {                                           /* :150-160 */
        double acc[3] = {0, 0, 0};
        for (int m = 0; m < n; ++m) {
          const double* C = c2w + 12 * (size_t)ms[m].frame;
          const double* c = ms[m].camera_coordinates;
          for (int r = 0; r < 3; ++r) acc[r] += ((C[4 * r] * c[0] + C[4 * r + 1] * c[1]) + C[4 * r + 2] * c[2]) + C[4 * r + 3];
        }
        for (int r = 0; r < 3; ++r) world[r] = acc[r] / n;
        outcome = 2;
      }
      ++it;
      break;
    }
    total_previous = total;                                                      /* :166 */
  }
  if (iterations) *iterations = (int)it;
  return outcome;
}

void orc_rotation_to_quaternion(const double R[9], double q[4]) {   /* Eigen 3.3 quaternionbase_assign_impl<Matrix3> */
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[7] - R[5]) * t;
    q[1] = (R[2] - R[6]) * t;
    q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (R[3 * k + j] - R[3 * j + k]) * t;
    q[j] = (R[3 * j + i] + R[3 * i + j]) * t;
    q[k] = (R[3 * k + i] + R[3 * i + k]) * t;
  }
}

int orc_format_trajectory_kitti(const double T[12], char* line, int capacity) {
  int n = 0;
  for (int u = 0; u < 3; ++u)
    for (int v = 0; v < 4; ++v) n += snprintf(line + n, n < capacity ? (size_t)(capacity - n) : 0, "%.9f ", T[4 * u + v]);
  n += snprintf(line + n, n < capacity ? (size_t)(capacity - n) : 0, "\n");
  return n;
}

int orc_format_trajectory_tum(double ts, const double T[12], char* line, int capacity) {
  const double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
  double q[4];
  orc_rotation_to_quaternion(R, q);
  return snprintf(line, (size_t)capacity, "%.9f %.9f %.9f %.9f %.9f %.9f %.9f %.9f \n", ts, T[3], T[7], T[11], q[0], q[1],
                  q[2], q[3]);
}
