// The change INTEGRATION.md section 2 makes to the reference's src/system/slam_assembly.cpp:62-76 and :82-96, as a
// translation unit of its own: type-checked against the reference's headers by tests/test_adapters_compile.py
// (g++ -std=c++14 -fsyntax-only; third-party headers are the shape-only stand-ins of tests/stubs/).
#include "position_tracking/pose_tracker_3d.h"
#include "types/parameters.h"

#include "gpu_frame_aligners.h"
#include "gpu_stereo_framepoint_generator.h"

namespace proslam {

void createStereoTrackerOnGpu(ParameterCollection* parameters_, PoseTracker3D* tracker_, Camera* camera_left_,
                              Camera* camera_right_, int cuda_device_) {
  StereoFramePointGenerator* framepoint_generator =
      new GpuStereoFramePointGenerator(parameters_->stereo_framepoint_generator_parameters, cuda_device_);
  framepoint_generator->setCameraLeft(camera_left_);
  framepoint_generator->setCameraRight(camera_right_);
  framepoint_generator->configure();

  StereoUVAligner* pose_optimizer = new GpuStereoUVAligner(parameters_->tracker_parameters->aligner, cuda_device_);
  pose_optimizer->setMaximumReliableDepthMeters(parameters_->stereo_framepoint_generator_parameters->maximum_reliable_depth_meters);
  pose_optimizer->setMinimumReliableDepthMeters(parameters_->stereo_framepoint_generator_parameters->minimum_depth_meters);
  pose_optimizer->configure();

  tracker_->setFramePointGenerator(framepoint_generator);   // PoseTracker3D deletes both (pose_tracker_3d.cpp:27-28)
  tracker_->setAligner(pose_optimizer);
  tracker_->configure();
}

void useDepthAlignerOnGpu(ParameterCollection* parameters_, PoseTracker3D* tracker_, int cuda_device_) {
  UVDAligner* pose_optimizer = new GpuUVDAligner(parameters_->tracker_parameters->aligner, cuda_device_);
  pose_optimizer->setMaximumReliableDepthMeters(parameters_->depth_framepoint_generator_parameters->maximum_reliable_depth_meters);
  pose_optimizer->setMinimumReliableDepthMeters(parameters_->depth_framepoint_generator_parameters->minimum_depth_meters);
  pose_optimizer->configure();
  tracker_->setAligner(pose_optimizer);
}

// what SLAMAssembly::printReport does with the generator (slam_assembly.cpp:690-719): the GPU class must survive the cast
double reportThroughTheReferencesCast(BaseFramePointGenerator* generator_) {
  StereoFramePointGenerator* stereo = dynamic_cast<StereoFramePointGenerator*>(generator_);
  return stereo ? stereo->meanTriangulationSuccessRatio() : 0.0;
}

}  // namespace proslam
