// gpu_stereo_framepoint_generator.h -- drop-in for proslam::StereoFramePointGenerator backed by libvslam_b200.so.
//
// Compiles ONLY inside the reference tree (needs its headers: Eigen, OpenCV, srrg_core); it is not built in this
// repository's image.  It derives from StereoFramePointGenerator because SLAMAssembly::printReport reaches the
// generator through dynamic_cast<StereoFramePointGenerator*> (reference src/system/slam_assembly.cpp:690-719).
//
// Overridden: configure(), initialize(), track(), compute(), recoverPoints() -> CUDA, through include/vslam_b200.h only.
// Nothing of the per-frame work runs the inherited CPU code; the host side only materialises the FramePoint objects
// the rest of the reference (tracker, landmarks, map) works on.
#pragma once
#include "framepoint_generation/stereo_framepoint_generator.h"
#include "vslam_b200.h"

namespace proslam {

class GpuStereoFramePointGenerator : public StereoFramePointGenerator {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  explicit GpuStereoFramePointGenerator(StereoFramePointGeneratorParameters* parameters_, int cuda_device_ = 0);
  ~GpuStereoFramePointGenerator() override;

  void configure() override;
  void initialize(Frame* frame_, const bool& extract_features_ = true) override;
  void track(Frame* frame_, Frame* frame_previous_, const TransformMatrix3D& camera_left_previous_in_current_,
             FramePointPointerVector& lost_points_, const bool track_by_appearance_ = true) override;
  void compute(Frame* frame_) override;
  void recoverPoints(Frame* current_frame_, const FramePointPointerVector& lost_points_) const override;

  // test table of cv::xfeatures2d::BriefDescriptorExtractor(32), needed before configure() when descriptor_type is
  // BRIEF in a build with opencv_contrib: 256 x (y0, x0, y1, x1) in the order of generated_32.i
  void setBriefTests(const int8_t tests_[1024]);

  // device seconds, the GPU counterparts of getTimeConsumptionSeconds_{keypoint_detection, ...}
  double deviceSecondsKeypointDetection() const;
  double deviceSecondsDescriptorExtraction() const;
  double deviceSecondsPointTriangulation() const;

 private:
  void refreshChronometers();
  void check(int status_) const;   // rethrows C-ABI errors as std::runtime_error (caught in executables/app.cpp:128)

  StereoFramePointGeneratorParameters* _stereo_parameters = nullptr;
  int _cuda_device = 0;
  vslam_fpg* _handle = nullptr;
  std::vector<vslam_keypoint> _keypoint_buffer;
  std::vector<vslam_framepoint> _framepoint_buffer;
  std::vector<vslam_tracked_point> _tracked_buffer;
  mutable std::vector<vslam_previous_point> _previous_buffer;
  std::vector<vslam_track> _track_buffer;
  std::vector<int32_t> _lost_buffer;
  mutable std::vector<vslam_recovered_point> _recovered_buffer;

  int8_t _brief_tests[1024] = {};
  bool _brief_tests_set = false;

  static void fillPreviousPoint(const FramePoint* point_, vslam_previous_point& out_);
};

}  // namespace proslam
