// gpu_stereo_framepoint_generator.h -- drop-in for proslam::StereoFramePointGenerator backed by libvslam_b200.so.
//
// Compiles ONLY inside the reference tree (needs its headers: Eigen, OpenCV, srrg_core); it is not built in this
// repository's image.  It derives from StereoFramePointGenerator because SLAMAssembly::printReport reaches the
// generator through dynamic_cast<StereoFramePointGenerator*> (reference src/system/slam_assembly.cpp:690-719).
//
// Overridden: configure(), initialize(), compute()      -> CUDA, through include/vslam_b200.h only
// Inherited : track(), recoverPoints()                   -> the reference's CPU code (SURVEY.md section 8(f) "next")
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
  void compute(Frame* frame_) override;

  // device seconds, the GPU counterparts of getTimeConsumptionSeconds_{keypoint_detection, ...}
  double deviceSecondsKeypointDetection() const;
  double deviceSecondsDescriptorExtraction() const;
  double deviceSecondsPointTriangulation() const;

 private:
  void check(int status_) const;   // rethrows C-ABI errors as std::runtime_error (caught in executables/app.cpp:128)

  StereoFramePointGeneratorParameters* _stereo_parameters = nullptr;
  int _cuda_device = 0;
  vslam_fpg* _handle = nullptr;
  std::vector<vslam_keypoint> _keypoint_buffer;
  std::vector<vslam_framepoint> _framepoint_buffer;
  std::vector<vslam_tracked_point> _tracked_buffer;
};

}  // namespace proslam
