// Shape-only stand-in for OpenCV (tests/stubs/README.md): declarations for type checking, nothing is ever linked.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>
#include "core/version.hpp"
#define CV_8U 0
#define CV_16U 2
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_16UC1 2
#define CV_32FC1 5
#define CV_64FC1 6
typedef unsigned char uchar;
namespace cv {
enum NormTypes { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_HAMMING = 6, NORM_HAMMING2 = 7 };
enum { RANSAC = 8, LMEDS = 4 };
template <class T> struct Point_ {
  Point_(); Point_(T, T);
  template <class U> Point_(const Point_<U>&);
  T x, y;
};
template <class T> Point_<T> operator-(const Point_<T>&, const Point_<T>&);
template <class T> Point_<T> operator+(const Point_<T>&, const Point_<T>&);
typedef Point_<int> Point; typedef Point_<int> Point2i; typedef Point_<float> Point2f; typedef Point_<double> Point2d;
template <class T> struct Rect_ {
  Rect_(); Rect_(T, T, T, T);
  T x, y, width, height;
};
typedef Rect_<int> Rect;
template <class T> struct Size_ { Size_(); Size_(T, T); T width, height; };
typedef Size_<int> Size;
template <class T, int N> struct Vec { Vec(); T val[N]; T& operator[](int); const T& operator()(int) const; };
struct Scalar { Scalar(); Scalar(double, double = 0, double = 0, double = 0); double val[4]; };
struct Range { Range(int, int); };
struct MatStep { operator size_t() const; size_t operator[](int) const; };
class Mat {
 public:
  Mat(); Mat(int, int, int); Mat(int, int, int, const Scalar&); Mat(int, int, int, void*, size_t = 0);
  Mat(const Mat&, const Rect&);
  Mat operator()(const Rect&) const;
  Mat row(int) const; Mat col(int) const; Mat rowRange(int, int) const; Mat clone() const;
  void create(int, int, int); void release(); void copyTo(Mat&) const; void convertTo(Mat&, int, double = 1, double = 0) const;
  bool empty() const; int type() const; int channels() const; int depth() const; size_t total() const; size_t elemSize() const;
  bool isContinuous() const; Size size() const;
  template <class T> T& at(int, int); template <class T> const T& at(int, int) const;
  template <class T> T& at(int); template <class T> const T& at(int) const;
  template <class T> T& at(Point); template <class T> const T& at(Point) const;
  template <class T> T* ptr(int = 0); template <class T> const T* ptr(int = 0) const;
  uchar* ptr(int = 0); const uchar* ptr(int = 0) const;
  void push_back(const Mat&);
  static Mat zeros(int, int, int); static Mat eye(int, int, int);
  int flags, dims, rows, cols;
  uchar* data;
  MatStep step;
};
typedef const Mat& InputArray; typedef Mat& OutputArray; typedef Mat& InputOutputArray;
struct KeyPoint {
  KeyPoint(); KeyPoint(Point2f, float, float = -1, float = 0, int = 0, int = -1);
  KeyPoint(float, float, float, float = -1, float = 0, int = 0, int = -1);
  Point2f pt; float size, angle, response; int octave, class_id;
};
struct DMatch { DMatch(); int queryIdx, trainIdx, imgIdx; float distance; };
template <class T> struct Ptr {
  Ptr(); Ptr(T*);
  template <class U> Ptr(const Ptr<U>&);
  T* operator->() const; T& operator*() const; T* get() const; operator T*() const; bool empty() const; void release();
};
class Feature2D {
 public:
  virtual ~Feature2D();
  virtual void detect(InputArray, std::vector<KeyPoint>&, InputArray = Mat());
  virtual void compute(InputArray, std::vector<KeyPoint>&, OutputArray);
  virtual void detectAndCompute(InputArray, InputArray, std::vector<KeyPoint>&, OutputArray, bool = false);
};
typedef Feature2D FeatureDetector; typedef Feature2D DescriptorExtractor;
class FastFeatureDetector : public Feature2D {
 public:
  enum { TYPE_5_8 = 0, TYPE_7_12 = 1, TYPE_9_16 = 2 };
  static Ptr<FastFeatureDetector> create(int = 10, bool = true, int = 2);
  virtual void setThreshold(int); virtual int getThreshold() const;
  virtual void setNonmaxSuppression(bool); virtual bool getNonmaxSuppression() const;
};
class AgastFeatureDetector : public Feature2D { public: static Ptr<AgastFeatureDetector> create(int = 10, bool = true, int = 3); virtual void setThreshold(int); virtual int getThreshold() const; };
class ORB : public Feature2D { public: static Ptr<ORB> create(int = 500, float = 1.2f, int = 8, int = 31, int = 0, int = 2, int = 0, int = 31, int = 20); virtual void setFastThreshold(int); virtual int getFastThreshold() const; };
class BRISK : public Feature2D { public: static Ptr<BRISK> create(int = 30, int = 3, float = 1.0f); };
class KAZE : public Feature2D { public: static Ptr<KAZE> create(bool = false, bool = false, float = 0.001f, int = 4, int = 4, int = 1); virtual void setThreshold(double); virtual double getThreshold() const; };
class AKAZE : public Feature2D { public: static Ptr<AKAZE> create(int = 5, int = 0, int = 3, float = 0.001f, int = 4, int = 4, int = 1); virtual void setThreshold(double); virtual double getThreshold() const; };
namespace xfeatures2d {
class SIFT : public Feature2D { public: static Ptr<SIFT> create(int = 0, int = 3, double = 0.04, double = 10, double = 1.6); };
class SURF : public Feature2D { public: static Ptr<SURF> create(double = 100); };
class FREAK : public Feature2D { public: static Ptr<FREAK> create(); };
class BriefDescriptorExtractor : public Feature2D { public: static Ptr<BriefDescriptorExtractor> create(int = 32, bool = false); };
}  // namespace xfeatures2d
class DescriptorMatcher {
 public:
  enum { FLANNBASED = 1, BRUTEFORCE = 2, BRUTEFORCE_L1 = 3, BRUTEFORCE_HAMMING = 4 };
  static Ptr<DescriptorMatcher> create(int); static Ptr<DescriptorMatcher> create(const std::string&);
  void match(InputArray, InputArray, std::vector<DMatch>&) const;
  void knnMatch(InputArray, InputArray, std::vector<std::vector<DMatch>>&, int) const;
};
class FlannBasedMatcher : public DescriptorMatcher { public: FlannBasedMatcher(); };
class BFMatcher : public DescriptorMatcher { public: BFMatcher(int = NORM_L2, bool = false); };
template <class T, int N> void Rodrigues(InputArray, Vec<T, N>&);
double norm(InputArray, InputArray, int = NORM_L2);
double norm(InputArray, int = NORM_L2);
Mat findHomography(const std::vector<Point2f>&, const std::vector<Point2f>&, int = 0, double = 3);
void imshow(const std::string&, InputArray); int waitKey(int = 0);
}  // namespace cv
