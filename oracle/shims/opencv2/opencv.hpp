// oracle/shims/opencv2/opencv.hpp -- FUNCTIONAL minimal stand-in for the part of OpenCV 3.x the reference's hot path
// uses.  TEST INFRASTRUCTURE ONLY (see oracle/shims/Eigen/Core): it lets the reference's unmodified translation units
// compile and run here, where OpenCV C++ is not installed.  It is not OpenCV and shares no code with it.
//
//   cv::Mat                reference-counted 2-D array with ROI views (u8 / u16 / f32 / f64, 1 or 3 channels)
//   cv::FastFeatureDetector::detect, cv::ORB::compute, cv::xfeatures2d::BriefDescriptorExtractor::compute
//                          call the backend installed through vslam_shim_set_backend() (oracle/shims/src/cv_shim.cpp):
//                          by default oracle/c/vslam_oracle.c (tier A, pinned bit for bit to cv2 4.13 by
//                          tests/test_oracle_vs_cv2.py); bench.py installs cv2's own FAST / ORB through ctypes callbacks
//   cv::norm(NORM_HAMMING) popcount
//   cv::Rodrigues          rotation matrix -> rotation vector (the formula of calib3d's cvRodrigues2 without its SVD
//                          re-orthonormalisation of the input)
//   the other detectors / extractors / matchers the reference names (AKAZE, KAZE, BRISK, AGAST, SIFT, FREAK, FLANN)
//                          are declared so that base_framepoint_generator.cpp compiles; using one throws.
//   DescriptorMatcher::knnMatch / findHomography: the reference's `use_matches` block
//                          (stereo_framepoint_generator.cpp:168-273) computes results nothing reads; here they return
//                          empty results, which leaves every output of compute() unchanged.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>
#include "core/version.hpp"

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
typedef unsigned char uchar;
typedef unsigned short ushort;

namespace cv {

enum NormTypes { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_L2SQR = 5, NORM_HAMMING = 6, NORM_HAMMING2 = 7 };
enum { LMEDS = 4, RANSAC = 8, RHO = 16 };

template <class T, class U> inline T saturate_cast(U v) { return static_cast<T>(v); }
template <> inline int saturate_cast<int, float>(float v) { return (int)std::lrint(v); }     // cvRound
template <> inline int saturate_cast<int, double>(double v) { return (int)std::lrint(v); }

template <class T>
struct Point_ {
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  template <class U> Point_(const Point_<U>& o) : x(saturate_cast<T>(o.x)), y(saturate_cast<T>(o.y)) {}
  Point_& operator+=(const Point_& o) { x += o.x; y += o.y; return *this; }
  Point_& operator-=(const Point_& o) { x -= o.x; y -= o.y; return *this; }
  T x, y;
};
template <class T> Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <class T> Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <class T> bool operator==(const Point_<T>& a, const Point_<T>& b) { return a.x == b.x && a.y == b.y; }
typedef Point_<int> Point; typedef Point_<int> Point2i; typedef Point_<float> Point2f; typedef Point_<double> Point2d;

template <class T>
struct Size_ {
  Size_() : width(0), height(0) {}
  Size_(T w, T h) : width(w), height(h) {}
  T area() const { return width * height; }
  T width, height;
};
typedef Size_<int> Size;
template <class T> std::ostream& operator<<(std::ostream& os, const Size_<T>& s) { return os << "[" << s.width << " x " << s.height << "]"; }

template <class T>
struct Rect_ {
  Rect_() : x(0), y(0), width(0), height(0) {}
  Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
  template <class U> operator Rect_<U>() const {
    return Rect_<U>(saturate_cast<U>(x), saturate_cast<U>(y), saturate_cast<U>(width), saturate_cast<U>(height));
  }
  Point_<T> tl() const { return Point_<T>(x, y); }
  Point_<T> br() const { return Point_<T>(x + width, y + height); }
  Size_<T> size() const { return Size_<T>(width, height); }
  T area() const { return width * height; }
  T x, y, width, height;
};
typedef Rect_<int> Rect; typedef Rect_<float> Rect2f; typedef Rect_<double> Rect2d;

template <class T, int N>
struct Vec {
  Vec() { for (int i = 0; i < N; ++i) val[i] = T(0); }
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
  T& operator()(int i) { return val[i]; }
  const T& operator()(int i) const { return val[i]; }
  T val[N];
};
typedef Vec<double, 3> Vec3d; typedef Vec<float, 3> Vec3f; typedef Vec<uchar, 3> Vec3b;

struct Scalar {
  Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
  double val[4];
};

struct MatStep {
  MatStep() { p[0] = p[1] = 0; }
  operator size_t() const { return p[0]; }
  size_t operator[](int i) const { return p[i]; }
  MatStep& operator=(size_t s) { p[0] = s; return *this; }
  size_t p[2];
};

class Mat {
 public:
  enum { AUTO_STEP = 0 };
  Mat() : flags(0), dims(0), rows(0), cols(0), data(nullptr) {}
  Mat(int rows_, int cols_, int type_) : Mat() { create(rows_, cols_, type_); }
  Mat(int rows_, int cols_, int type_, const Scalar& s);
  Mat(int rows_, int cols_, int type_, void* data_, size_t step_ = AUTO_STEP);   // external memory, not owned
  Mat(const Mat& m, const Rect& roi);
  Mat operator()(const Rect& roi) const { return Mat(*this, roi); }
  Mat row(int y) const { return Mat(*this, Rect(0, y, cols, 1)); }
  Mat col(int x) const { return Mat(*this, Rect(x, 0, 1, rows)); }
  Mat rowRange(int y0, int y1) const { return Mat(*this, Rect(0, y0, cols, y1 - y0)); }
  Mat clone() const;
  void create(int rows_, int cols_, int type_);
  void release();
  void copyTo(Mat& dst) const;
  void convertTo(Mat& dst, int rtype, double alpha = 1, double beta = 0) const;
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  int type() const { return flags & 0xfff; }
  int depth() const { return flags & 7; }
  int channels() const { return ((flags & 0xfff) >> CV_CN_SHIFT) + 1; }
  size_t elemSize1() const { static const int s[8] = {1, 1, 2, 2, 4, 4, 8, 0}; return s[depth()]; }
  size_t elemSize() const { return elemSize1() * channels(); }
  size_t total() const { return (size_t)rows * cols; }
  bool isContinuous() const { return rows <= 1 || step.p[0] == (size_t)cols * elemSize(); }
  Size size() const { return Size(cols, rows); }
  uchar* ptr(int y = 0) { return data + (size_t)y * step.p[0]; }
  const uchar* ptr(int y = 0) const { return data + (size_t)y * step.p[0]; }
  template <class T> T* ptr(int y = 0) { return reinterpret_cast<T*>(ptr(y)); }
  template <class T> const T* ptr(int y = 0) const { return reinterpret_cast<const T*>(ptr(y)); }
  template <class T> T& at(int y, int x) { return ptr<T>(y)[x]; }
  template <class T> const T& at(int y, int x) const { return ptr<T>(y)[x]; }
  template <class T> T& at(int i) { return rows == 1 ? ptr<T>(0)[i] : ptr<T>(i)[0]; }
  template <class T> const T& at(int i) const { return rows == 1 ? ptr<T>(0)[i] : ptr<T>(i)[0]; }
  template <class T> T& at(Point p) { return ptr<T>(p.y)[p.x]; }
  template <class T> const T& at(Point p) const { return ptr<T>(p.y)[p.x]; }
  static Mat zeros(int rows_, int cols_, int type_);
  static Mat eye(int rows_, int cols_, int type_);

  int flags, dims, rows, cols;
  uchar* data;
  MatStep step;
 private:
  std::shared_ptr<std::vector<uchar>> _owner;   // keeps the allocation alive for every header sharing it
};
typedef const Mat& InputArray;
typedef Mat& OutputArray;
typedef Mat& InputOutputArray;
inline const Mat& noArray() { static const Mat none; return none; }

struct KeyPoint {
  KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(Point2f pt_, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
      : pt(pt_), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
  KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
      : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
};

struct DMatch {
  DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(0) {}
  int queryIdx, trainIdx, imgIdx;
  float distance;
};

template <class T>
class Ptr {
 public:
  Ptr() {}
  Ptr(T* p) : _p(p) {}
  Ptr(const std::shared_ptr<T>& p) : _p(p) {}
  template <class U> Ptr(const Ptr<U>& o) : _p(o.shared()) {}
  T* operator->() const { return _p.get(); }
  T& operator*() const { return *_p; }
  T* get() const { return _p.get(); }
  operator T*() const { return _p.get(); }
  bool empty() const { return !_p; }
  void release() { _p.reset(); }
  const std::shared_ptr<T>& shared() const { return _p; }
 private:
  std::shared_ptr<T> _p;
};
template <class T, class... A> Ptr<T> makePtr(A&&... a) { return Ptr<T>(std::make_shared<T>(std::forward<A>(a)...)); }

class Feature2D {
 public:
  virtual ~Feature2D() {}
  virtual void detect(InputArray image, std::vector<KeyPoint>& keypoints, InputArray mask = noArray());
  virtual void compute(InputArray image, std::vector<KeyPoint>& keypoints, OutputArray descriptors);
  virtual void detectAndCompute(InputArray image, InputArray mask, std::vector<KeyPoint>& keypoints,
                                OutputArray descriptors, bool useProvidedKeypoints = false);
 protected:
  [[noreturn]] void unavailable(const char* what) const;
};
typedef Feature2D FeatureDetector;
typedef Feature2D DescriptorExtractor;

class FastFeatureDetector : public Feature2D {
 public:
  enum { TYPE_5_8 = 0, TYPE_7_12 = 1, TYPE_9_16 = 2 };
  static Ptr<FastFeatureDetector> create(int threshold = 10, bool nonmaxSuppression = true, int type = TYPE_9_16);
  FastFeatureDetector(int threshold, bool nms, int type) : _threshold(threshold), _nms(nms), _type(type) {}
  virtual void setThreshold(int threshold) { _threshold = threshold; }
  virtual int getThreshold() const { return _threshold; }
  virtual void setNonmaxSuppression(bool f) { _nms = f; }
  virtual bool getNonmaxSuppression() const { return _nms; }
  virtual int getType() const { return _type; }
  void detect(InputArray image, std::vector<KeyPoint>& keypoints, InputArray mask = noArray()) override;
 private:
  int _threshold;
  bool _nms;
  int _type;
};

class AgastFeatureDetector : public Feature2D {
 public:
  static Ptr<AgastFeatureDetector> create(int threshold = 10, bool nonmaxSuppression = true, int type = 3);
  explicit AgastFeatureDetector(int t) : _threshold(t) {}
  virtual void setThreshold(int t) { _threshold = t; }
  virtual int getThreshold() const { return _threshold; }
 private:
  int _threshold;
};

class ORB : public Feature2D {
 public:
  enum { kBytes = 32, HARRIS_SCORE = 0, FAST_SCORE = 1 };
  static Ptr<ORB> create(int nfeatures = 500, float scaleFactor = 1.2f, int nlevels = 8, int edgeThreshold = 31,
                         int firstLevel = 0, int WTA_K = 2, int scoreType = HARRIS_SCORE, int patchSize = 31,
                         int fastThreshold = 20);
  ORB(int edge_threshold, int fast_threshold) : _edge_threshold(edge_threshold), _fast_threshold(fast_threshold) {}
  virtual void setFastThreshold(int t) { _fast_threshold = t; }
  virtual int getFastThreshold() const { return _fast_threshold; }
  virtual void setEdgeThreshold(int t) { _edge_threshold = t; }
  virtual int getEdgeThreshold() const { return _edge_threshold; }
  void compute(InputArray image, std::vector<KeyPoint>& keypoints, OutputArray descriptors) override;
 private:
  int _edge_threshold, _fast_threshold;
};

class BRISK : public Feature2D {
 public:
  static Ptr<BRISK> create(int thresh = 30, int octaves = 3, float patternScale = 1.0f);
};

class KAZE : public Feature2D {
 public:
  enum { DIFF_PM_G1 = 0, DIFF_PM_G2 = 1, DIFF_WEICKERT = 2, DIFF_CHARBONNIER = 3 };
  static Ptr<KAZE> create(bool extended = false, bool upright = false, float threshold = 0.001f, int nOctaves = 4,
                          int nOctaveLayers = 4, int diffusivity = DIFF_PM_G2);
  explicit KAZE(double t) : _threshold(t) {}
  virtual void setThreshold(double t) { _threshold = t; }
  virtual double getThreshold() const { return _threshold; }
 private:
  double _threshold;
};

class AKAZE : public Feature2D {
 public:
  enum { DESCRIPTOR_KAZE_UPRIGHT = 2, DESCRIPTOR_KAZE = 3, DESCRIPTOR_MLDB_UPRIGHT = 4, DESCRIPTOR_MLDB = 5 };
  static Ptr<AKAZE> create(int descriptor_type = DESCRIPTOR_MLDB, int descriptor_size = 0, int descriptor_channels = 3,
                           float threshold = 0.001f, int nOctaves = 4, int nOctaveLayers = 4,
                           int diffusivity = KAZE::DIFF_PM_G2);
  explicit AKAZE(double t) : _threshold(t) {}
  virtual void setThreshold(double t) { _threshold = t; }
  virtual double getThreshold() const { return _threshold; }
 private:
  double _threshold;
};

namespace xfeatures2d {
class SIFT : public Feature2D {
 public:
  static Ptr<SIFT> create(int nfeatures = 0, int nOctaveLayers = 3, double contrastThreshold = 0.04,
                          double edgeThreshold = 10, double sigma = 1.6);
};
class SURF : public Feature2D {
 public:
  static Ptr<SURF> create(double hessianThreshold = 100);
};
class FREAK : public Feature2D {
 public:
  static Ptr<FREAK> create();
};
class BriefDescriptorExtractor : public Feature2D {
 public:
  static Ptr<BriefDescriptorExtractor> create(int bytes = 32, bool use_orientation = false);
  explicit BriefDescriptorExtractor(int bytes) : _bytes(bytes) {}
  void compute(InputArray image, std::vector<KeyPoint>& keypoints, OutputArray descriptors) override;
 private:
  int _bytes;
};
}  // namespace xfeatures2d

class DescriptorMatcher {
 public:
  enum { FLANNBASED = 1, BRUTEFORCE = 2, BRUTEFORCE_L1 = 3, BRUTEFORCE_HAMMING = 4, BRUTEFORCE_HAMMINGLUT = 5,
         BRUTEFORCE_SL2 = 6 };
  virtual ~DescriptorMatcher() {}
  static Ptr<DescriptorMatcher> create(int matcherType);
  static Ptr<DescriptorMatcher> create(const std::string& descriptorMatcherType);
  // see the header comment: empty results (the reference never reads them)
  void match(InputArray query, InputArray train, std::vector<DMatch>& matches) const;
  void knnMatch(InputArray query, InputArray train, std::vector<std::vector<DMatch>>& matches, int k) const;
};

double norm(InputArray a, InputArray b, int normType = NORM_L2);
double norm(InputArray a, int normType = NORM_L2);
template <class T> inline double norm(const Point_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }

void Rodrigues(InputArray rotation_matrix, double rotation_vector[3]);
template <class T, int N>
inline void Rodrigues(InputArray rotation_matrix, Vec<T, N>& rotation_vector) {
  static_assert(N == 3, "rotation vector");
  double r[3];
  Rodrigues(rotation_matrix, r);
  for (int i = 0; i < 3; ++i) rotation_vector[i] = static_cast<T>(r[i]);
}

template <class P>
inline Mat findHomography(const std::vector<P>&, const std::vector<P>&, int = 0, double = 3, OutputArray mask = const_cast<Mat&>(noArray()),
                          int = 2000, double = 0.995) {
  (void)mask;
  return Mat();
}

inline void setNumThreads(int) {}
inline void setUseOptimized(bool) {}

}  // namespace cv

// ---- backend of the three primitives (C linkage: Python installs cv2-backed callbacks through ctypes) --------------
extern "C" {
// keypoints: x, y, response triples (FAST: size 7, angle -1, octave 0, class_id -1 are implied)
typedef int (*vslam_shim_fast_fn)(const uint8_t* image, int stride, int cols, int rows, int threshold, float* xyr, int capacity);
// filters keypoints in place (order kept), writes n x 32 descriptor bytes, returns n
typedef int (*vslam_shim_describe_fn)(const uint8_t* image, int stride, int cols, int rows, float* xyr, int n, uint8_t* descriptors);
void vslam_shim_set_backend(vslam_shim_fast_fn fast, vslam_shim_describe_fn orb, vslam_shim_describe_fn brief);
void vslam_shim_set_brief_tests(const int8_t tests[1024]);
}
