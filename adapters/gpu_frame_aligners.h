// gpu_frame_aligners.h -- drop-ins for proslam::StereoUVAligner / proslam::UVDAligner backed by libvslam_b200.so.
// Compiles only inside the reference tree.  initialize() stays the reference's host code (it walks the Frame /
// FramePoint / Landmark object graph) and is followed by ONE upload; linearize / oneRound /
// converge run through include/vslam_b200.h.
#pragma once
#include "aligners/stereouv_aligner.h"
#include "aligners/uvd_aligner.h"
#include "vslam_b200.h"

namespace proslam {

template <class ReferenceAligner, int Kind>
class GpuFrameAligner : public ReferenceAligner {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  explicit GpuFrameAligner(AlignerParameters* parameters_, int cuda_device_ = 0);
  ~GpuFrameAligner() override;

  void initialize(const Frame* frame_previous_, const Frame* frame_current_,
                  const TransformMatrix3D& previous_to_current_) override;
  void linearize(const bool& ignore_outliers_) override;
  void oneRound(const bool& ignore_outliers_) override;
  void converge() override;

 private:
  void check(int status_) const;
  void adopt(const vslam_linear_system& system_);     // _H, _b, _total_error, _number_of_inliers / _outliers
  void pose_to_array(double T_[12]) const;
  void array_to_pose(const double T_[12]);
  vslam_aligner_parameters parameters_for_abi() const;
  void fetch_errors_and_inliers();

  int _cuda_device = 0;
  int32_t _capacity = 0;
  vslam_aligner* _handle = nullptr;
  std::vector<double> _pack;
  std::vector<uint8_t> _inlier_bytes;
};

typedef GpuFrameAligner<StereoUVAligner, VSLAM_ALIGNER_STEREO_UV> GpuStereoUVAligner;
typedef GpuFrameAligner<UVDAligner, VSLAM_ALIGNER_UVD> GpuUVDAligner;

}  // namespace proslam
