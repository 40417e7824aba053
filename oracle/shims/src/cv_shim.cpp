// oracle/shims/src/cv_shim.cpp -- implementation of the OpenCV stand-in (oracle/shims/opencv2/opencv.hpp).
// TEST INFRASTRUCTURE ONLY.  The three primitives the reference delegates to OpenCV are answered by a pluggable
// backend; the default is the tier-A oracle (oracle/c/vslam_oracle.c), whose FAST / ORB are pinned bit for bit to
// OpenCV 4.13 (tests/test_oracle_vs_cv2.py, tests/golden/).
#include <opencv2/opencv.hpp>

#include <algorithm>

#include "../../c/vslam_oracle.h"

namespace {

int default_fast(const uint8_t* image, int stride, int cols, int rows, int threshold, float* xyr, int capacity) {
  // cv::FastFeatureDetector::detect on a cols x rows view (base_framepoint_generator.cpp:23-25, :367)
  static_assert(sizeof(orc_kp) == 3 * sizeof(float), "orc_kp is x, y, response");
  return orc_fast_detect(image, stride, cols, rows, threshold, reinterpret_cast<orc_kp*>(xyr), capacity);
}
int default_orb(const uint8_t* image, int stride, int cols, int rows, float* xyr, int n, uint8_t* descriptors) {
  return orc_orb_compute(image, stride, cols, rows, nullptr, 0, reinterpret_cast<orc_kp*>(xyr), n, descriptors);
}
int8_t g_brief_tests[1024];
bool g_brief_tests_set = false;
int default_brief(const uint8_t* image, int stride, int cols, int rows, float* xyr, int n, uint8_t* descriptors) {
  if (!g_brief_tests_set)
    throw std::runtime_error("oracle/shims: BriefDescriptorExtractor needs vslam_shim_set_brief_tests() (opencv_contrib's "
                             "generated_32.i is not in this image)");
  return orc_brief32_compute(image, stride, cols, rows, g_brief_tests, reinterpret_cast<orc_kp*>(xyr), n, descriptors);
}

vslam_shim_fast_fn g_fast = default_fast;
vslam_shim_describe_fn g_orb = default_orb;
vslam_shim_describe_fn g_brief = default_brief;

const cv::Mat& require_u8(const cv::Mat& image, const char* who) {
  if (image.type() != CV_8UC1) throw std::runtime_error(std::string("oracle/shims: ") + who + " expects a CV_8UC1 image");
  return image;
}

void describe(vslam_shim_describe_fn fn, const cv::Mat& image, std::vector<cv::KeyPoint>& keypoints, cv::Mat& descriptors,
              const char* who) {
  require_u8(image, who);
  const int n = (int)keypoints.size();
  std::vector<float> xyr(3 * (size_t)std::max(n, 1));
  for (int i = 0; i < n; ++i) {
    xyr[3 * i] = keypoints[i].pt.x;
    xyr[3 * i + 1] = keypoints[i].pt.y;
    xyr[3 * i + 2] = (float)i;          // the backend keeps the order: carry the index to keep the other fields
  }
  std::vector<uint8_t> desc(32 * (size_t)std::max(n, 1));
  const int kept = n ? fn(image.data, (int)image.step, image.cols, image.rows, xyr.data(), n, desc.data()) : 0;
  std::vector<cv::KeyPoint> out(kept);
  for (int i = 0; i < kept; ++i) out[i] = keypoints[(int)xyr[3 * i + 2]];
  keypoints.swap(out);
  if (kept == 0) {
    descriptors.release();            // cv::ORB::compute releases the output when nothing is left
    return;
  }
  descriptors.create(kept, 32, CV_8UC1);
  for (int i = 0; i < kept; ++i) std::memcpy(descriptors.ptr(i), desc.data() + 32 * (size_t)i, 32);
}

}  // namespace

extern "C" void vslam_shim_set_backend(vslam_shim_fast_fn fast, vslam_shim_describe_fn orb, vslam_shim_describe_fn brief) {
  g_fast = fast ? fast : default_fast;
  g_orb = orb ? orb : default_orb;
  g_brief = brief ? brief : default_brief;
}
extern "C" void vslam_shim_set_brief_tests(const int8_t tests[1024]) {
  std::memcpy(g_brief_tests, tests, sizeof(g_brief_tests));
  g_brief_tests_set = true;
}

namespace cv {

// ---- Mat -----------------------------------------------------------------------------------------------------------
Mat::Mat(int rows_, int cols_, int type_, const Scalar& s) : Mat() {
  create(rows_, cols_, type_);
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols * channels(); ++x) {
      const double v = s.val[x % channels()];
      switch (depth()) {
        case CV_8U: ptr<uchar>(y)[x] = (uchar)v; break;
        case CV_16U: ptr<ushort>(y)[x] = (ushort)v; break;
        case CV_32S: ptr<int>(y)[x] = (int)v; break;
        case CV_32F: ptr<float>(y)[x] = (float)v; break;
        case CV_64F: ptr<double>(y)[x] = v; break;
        default: throw std::runtime_error("oracle/shims: Mat depth not supported");
      }
    }
}

Mat::Mat(int rows_, int cols_, int type_, void* data_, size_t step_) : Mat() {
  flags = type_ & 0xfff;
  dims = 2;
  rows = rows_;
  cols = cols_;
  data = static_cast<uchar*>(data_);
  step.p[1] = elemSize();
  step.p[0] = step_ == AUTO_STEP ? (size_t)cols * elemSize() : step_;
}

Mat::Mat(const Mat& m, const Rect& roi) : Mat() {
  if (roi.x < 0 || roi.y < 0 || roi.width < 0 || roi.height < 0 || roi.x + roi.width > m.cols || roi.y + roi.height > m.rows)
    throw std::runtime_error("oracle/shims: Mat ROI outside the matrix (OpenCV raises an assertion here)");
  flags = m.flags;
  dims = 2;
  rows = roi.height;
  cols = roi.width;
  step = m.step;
  data = m.data + (size_t)roi.y * m.step.p[0] + (size_t)roi.x * m.elemSize();
  _owner = m._owner;
}

void Mat::create(int rows_, int cols_, int type_) {
  if (data && rows == rows_ && cols == cols_ && type() == (type_ & 0xfff) && _owner && isContinuous()) return;
  flags = type_ & 0xfff;
  dims = 2;
  rows = rows_;
  cols = cols_;
  step.p[1] = elemSize();
  step.p[0] = (size_t)cols * elemSize();
  _owner = std::make_shared<std::vector<uchar>>((size_t)rows * step.p[0] + 64);
  data = _owner->data();
}

void Mat::release() {
  _owner.reset();
  data = nullptr;
  rows = cols = 0;
}

Mat Mat::clone() const {
  Mat m;
  copyTo(m);
  return m;
}

void Mat::copyTo(Mat& dst) const {
  if (empty()) {
    dst.release();
    return;
  }
  dst.create(rows, cols, type());
  for (int y = 0; y < rows; ++y) std::memcpy(dst.ptr(y), ptr(y), (size_t)cols * elemSize());
}

void Mat::convertTo(Mat& dst, int rtype, double alpha, double beta) const {
  Mat out;
  out.create(rows, cols, CV_MAKETYPE(rtype & 7, channels()));
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols * channels(); ++x) {
      double v;
      switch (depth()) {
        case CV_8U: v = ptr<uchar>(y)[x]; break;
        case CV_16U: v = ptr<ushort>(y)[x]; break;
        case CV_32S: v = ptr<int>(y)[x]; break;
        case CV_32F: v = ptr<float>(y)[x]; break;
        case CV_64F: v = ptr<double>(y)[x]; break;
        default: throw std::runtime_error("oracle/shims: Mat depth not supported");
      }
      v = v * alpha + beta;
      switch (out.depth()) {
        case CV_8U: out.ptr<uchar>(y)[x] = (uchar)std::min(255.0, std::max(0.0, std::nearbyint(v))); break;
        case CV_16U: out.ptr<ushort>(y)[x] = (ushort)std::min(65535.0, std::max(0.0, std::nearbyint(v))); break;
        case CV_32S: out.ptr<int>(y)[x] = (int)std::nearbyint(v); break;
        case CV_32F: out.ptr<float>(y)[x] = (float)v; break;
        case CV_64F: out.ptr<double>(y)[x] = v; break;
        default: throw std::runtime_error("oracle/shims: Mat depth not supported");
      }
    }
  dst = out;
}

Mat Mat::zeros(int rows_, int cols_, int type_) {
  Mat m(rows_, cols_, type_);
  for (int y = 0; y < m.rows; ++y) std::memset(m.ptr(y), 0, (size_t)m.cols * m.elemSize());
  return m;
}

Mat Mat::eye(int rows_, int cols_, int type_) {
  Mat m = zeros(rows_, cols_, type_);
  for (int i = 0; i < std::min(rows_, cols_); ++i) {
    if (m.depth() == CV_64F) m.at<double>(i, i) = 1;
    else if (m.depth() == CV_32F) m.at<float>(i, i) = 1;
    else if (m.depth() == CV_8U) m.at<uchar>(i, i) = 1;
  }
  return m;
}

// ---- features -----------------------------------------------------------------------------------------------------
void Feature2D::unavailable(const char* what) const {
  throw std::runtime_error(std::string("oracle/shims: ") + what + " is not available in the OpenCV stand-in (only "
                           "FastFeatureDetector::detect, ORB::compute and BriefDescriptorExtractor::compute are)");
}
void Feature2D::detect(InputArray, std::vector<KeyPoint>&, InputArray) { unavailable("this detector"); }
void Feature2D::compute(InputArray, std::vector<KeyPoint>&, OutputArray) { unavailable("this descriptor extractor"); }
void Feature2D::detectAndCompute(InputArray, InputArray, std::vector<KeyPoint>&, OutputArray, bool) { unavailable("detectAndCompute"); }

Ptr<FastFeatureDetector> FastFeatureDetector::create(int threshold, bool nms, int type) {
  return makePtr<FastFeatureDetector>(threshold, nms, type);
}

void FastFeatureDetector::detect(InputArray image, std::vector<KeyPoint>& keypoints, InputArray) {
  require_u8(image, "FastFeatureDetector::detect");
  if (!_nms || _type != TYPE_9_16) unavailable("FAST without non-maximum suppression or other than TYPE_9_16");
  keypoints.clear();
  if (image.empty()) return;
  const int threshold = std::min(std::max(_threshold, 0), 255);   // cv::FAST clamps the threshold to [0, 255]
  std::vector<float> xyr(3 * 4096);
  int n = g_fast(image.data, (int)image.step, image.cols, image.rows, threshold, xyr.data(), (int)xyr.size() / 3);
  if (n > (int)xyr.size() / 3) {
    xyr.resize(3 * (size_t)n);
    n = g_fast(image.data, (int)image.step, image.cols, image.rows, threshold, xyr.data(), n);
  }
  keypoints.resize(n);
  for (int i = 0; i < n; ++i) keypoints[i] = KeyPoint(xyr[3 * i], xyr[3 * i + 1], 7.f, -1.f, xyr[3 * i + 2], 0, -1);
}

Ptr<AgastFeatureDetector> AgastFeatureDetector::create(int threshold, bool, int) { return makePtr<AgastFeatureDetector>(threshold); }

Ptr<ORB> ORB::create(int, float, int, int edgeThreshold, int, int, int, int, int fastThreshold) {
  return makePtr<ORB>(edgeThreshold, fastThreshold);
}
void ORB::compute(InputArray image, std::vector<KeyPoint>& keypoints, OutputArray descriptors) {
  if (_edge_threshold != 31) unavailable("ORB with an edge threshold other than 31");
  // single-keypoint calls on a small region of interest (recoverPoints, stereo_framepoint_generator.cpp:769-795) stay on the
  // built-in backend: an installed callback (python cv2) would add more call overhead than the work itself
  const bool small = (long)image.rows * image.cols <= 128L * 128L;
  describe(small ? default_orb : g_orb, image, keypoints, descriptors, "ORB::compute");
}

Ptr<BRISK> BRISK::create(int, int, float) { return makePtr<BRISK>(); }
Ptr<KAZE> KAZE::create(bool, bool, float threshold, int, int, int) { return makePtr<KAZE>((double)threshold); }
Ptr<AKAZE> AKAZE::create(int, int, int, float threshold, int, int, int) { return makePtr<AKAZE>((double)threshold); }

namespace xfeatures2d {
Ptr<SIFT> SIFT::create(int, int, double, double, double) { return makePtr<SIFT>(); }
Ptr<SURF> SURF::create(double) { return makePtr<SURF>(); }
Ptr<FREAK> FREAK::create() { return makePtr<FREAK>(); }
Ptr<BriefDescriptorExtractor> BriefDescriptorExtractor::create(int bytes, bool) { return makePtr<BriefDescriptorExtractor>(bytes); }
void BriefDescriptorExtractor::compute(InputArray image, std::vector<KeyPoint>& keypoints, OutputArray descriptors) {
  if (_bytes != 32) unavailable("BRIEF with a size other than 32 bytes");
  describe(g_brief, image, keypoints, descriptors, "BriefDescriptorExtractor::compute");
}
}  // namespace xfeatures2d

Ptr<DescriptorMatcher> DescriptorMatcher::create(int) { return makePtr<DescriptorMatcher>(); }
Ptr<DescriptorMatcher> DescriptorMatcher::create(const std::string&) { return makePtr<DescriptorMatcher>(); }
void DescriptorMatcher::match(InputArray, InputArray, std::vector<DMatch>& matches) const { matches.clear(); }
void DescriptorMatcher::knnMatch(InputArray, InputArray, std::vector<std::vector<DMatch>>& matches, int) const { matches.clear(); }

// ---- norms ---------------------------------------------------------------------------------------------------------
double norm(InputArray a, InputArray b, int normType) {
  if (a.rows != b.rows || a.cols != b.cols || a.type() != b.type())
    throw std::runtime_error("oracle/shims: norm of arrays of different size or type");
  if (normType == NORM_HAMMING) {
    if (a.depth() != CV_8U) throw std::runtime_error("oracle/shims: NORM_HAMMING needs CV_8U");
    unsigned d = 0;
    const size_t w = (size_t)a.cols * a.channels();
    for (int y = 0; y < a.rows; ++y) {
      const uchar *p = a.ptr(y), *q = b.ptr(y);
      for (size_t x = 0; x < w; ++x) d += (unsigned)__builtin_popcount((unsigned)(p[x] ^ q[x]));
    }
    return d;
  }
  if (normType == NORM_L2 || normType == NORM_L2SQR || normType == NORM_L1 || normType == NORM_INF) {
    Mat da, db;
    a.convertTo(da, CV_64F);
    b.convertTo(db, CV_64F);
    double s = 0;
    for (int y = 0; y < da.rows; ++y)
      for (int x = 0; x < da.cols * da.channels(); ++x) {
        const double v = std::fabs(da.ptr<double>(y)[x] - db.ptr<double>(y)[x]);
        if (normType == NORM_L1) s += v;
        else if (normType == NORM_INF) s = std::max(s, v);
        else s += v * v;
      }
    return normType == NORM_L2 ? std::sqrt(s) : s;
  }
  throw std::runtime_error("oracle/shims: norm type not supported");
}

double norm(InputArray a, int normType) {
  Mat zero = Mat::zeros(a.rows, a.cols, a.type());
  return norm(a, zero, normType);
}

// ---- Rodrigues: rotation matrix -> rotation vector -------------------------------------------------------------------
// The formula of OpenCV's cvRodrigues2 for a 3 x 3 input, WITHOUT its first step (an SVD that replaces R by the nearest
// rotation): the reference only passes rotations the aligner has just re-orthonormalised (stereouv_aligner.cpp:203-206)
// and compares the norm of the result with thresholds (pose_tracker_3d.cpp:148, :377; world_map.cpp:113).
void Rodrigues(InputArray m, double r[3]) {
  if (m.rows != 3 || m.cols != 3 || m.type() != CV_64FC1) throw std::runtime_error("oracle/shims: Rodrigues expects a 3 x 3 CV_64F matrix");
  double R[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[3 * i + j] = m.at<double>(i, j);
  double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
  const double s = std::sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
  double c = (R[0] + R[4] + R[8] - 1) * 0.5;
  c = c > 1. ? 1. : c < -1. ? -1. : c;
  double theta = std::acos(c);
  if (s < 1e-5) {
    if (c > 0) {
      rx = ry = rz = 0;
    } else {
      double t = (R[0] + 1) * 0.5;
      rx = std::sqrt(std::max(t, 0.));
      t = (R[4] + 1) * 0.5;
      ry = std::sqrt(std::max(t, 0.)) * (R[1] < 0 ? -1. : 1.);
      t = (R[8] + 1) * 0.5;
      rz = std::sqrt(std::max(t, 0.)) * (R[2] < 0 ? -1. : 1.);
      if (std::fabs(rx) < std::fabs(ry) && std::fabs(rx) < std::fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
      theta /= std::sqrt(rx * rx + ry * ry + rz * rz);
      rx *= theta;
      ry *= theta;
      rz *= theta;
    }
  } else {
    double vth = 1 / (2 * s);
    vth *= theta;
    rx *= vth;
    ry *= vth;
    rz *= vth;
  }
  r[0] = rx;
  r[1] = ry;
  r[2] = rz;
}

}  // namespace cv
