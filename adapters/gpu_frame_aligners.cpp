// gpu_frame_aligners.cpp -- see the header.  Line citations: reference src/aligners/stereouv_aligner.cpp (S) and
// src/aligners/uvd_aligner.cpp (U).
#include "gpu_frame_aligners.h"

#include <stdexcept>

namespace proslam {

template <class A, int K>
GpuFrameAligner<A, K>::GpuFrameAligner(AlignerParameters* parameters_, int cuda_device_)
    : A(parameters_), _cuda_device(cuda_device_) {}

template <class A, int K>
GpuFrameAligner<A, K>::~GpuFrameAligner() { vslam_aligner_destroy(_handle); }

template <class A, int K>
void GpuFrameAligner<A, K>::check(int status_) const {
  if (status_ != VSLAM_OK) throw std::runtime_error(std::string("GpuFrameAligner|") + vslam_last_error());
}

template <class A, int K>
void GpuFrameAligner<A, K>::pose_to_array(double T_[12]) const {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) T_[4 * r + c] = this->_previous_to_current.matrix()(r, c);
}

template <class A, int K>
void GpuFrameAligner<A, K>::array_to_pose(const double T_[12]) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) this->_previous_to_current.matrix()(r, c) = T_[4 * r + c];
}

template <class A, int K>
vslam_aligner_parameters GpuFrameAligner<A, K>::parameters_for_abi() const {
  const AlignerParameters* p = this->_parameters;
  vslam_aligner_parameters q;
  q.error_delta_for_convergence = p->error_delta_for_convergence;
  q.maximum_error_kernel = p->maximum_error_kernel;
  q.damping = p->damping;
  q.maximum_number_of_iterations = (int32_t)p->maximum_number_of_iterations;
  q.minimum_number_of_inliers = (int32_t)p->minimum_number_of_inliers;
  return q;
}

template <class A, int K>
void GpuFrameAligner<A, K>::adopt(const vslam_linear_system& s) {
  for (int r = 0; r < 6; ++r) {
    for (int c = 0; c < 6; ++c) this->_H(r, c) = s.H[6 * r + c];
    this->_b(r) = s.b[r];
  }
  this->_total_error = s.total_error;
  this->_number_of_inliers = s.number_of_inliers;
  this->_number_of_outliers = s.number_of_outliers;
}

template <class A, int K>
void GpuFrameAligner<A, K>::initialize(const Frame* frame_previous_, const Frame* frame_current_,
                                       const TransformMatrix3D& previous_to_current_) {
  A::initialize(frame_previous_, frame_current_, previous_to_current_);   // S:10-69 / U:11-74, host object graph walk
  const int32_t n = (int32_t)this->_number_of_measurements;
  if (!_handle || n > _capacity) {
    vslam_aligner_destroy(_handle);
    _handle = nullptr;
    _capacity = n > 4096 ? 2 * n : 8192;
    check(vslam_aligner_create(K, _capacity, _cuda_device, &_handle));
  }
  constexpr int F = K == VSLAM_ALIGNER_STEREO_UV ? 4 : 3;   // measurement dimension
  constexpr int W = K == VSLAM_ALIGNER_STEREO_UV ? 1 : 2;   // information scalars per point
  _pack.resize((size_t)n * (3 + F + W + 1));
  double* moving = _pack.data();
  double* fixed = moving + 3 * (size_t)n;
  double* omega = fixed + F * (size_t)n;
  double* wt = omega + W * (size_t)n;
  for (int32_t u = 0; u < n; ++u) {
    for (int k = 0; k < 3; ++k) moving[3 * u + k] = this->_moving[u](k);
    for (int k = 0; k < F; ++k) fixed[F * u + k] = this->_fixed[u](k);
    omega[W * u] = this->_information_matrix_vector[u](0, 0);              // scalar * I (S:29,47) / diag(w, w, wd)
    if (W == 2) omega[W * u + 1] = this->_information_matrix_vector[u](2, 2);   // U:52-61
    wt[u] = this->_weights_translation[u];
  }
  double Kmat[9], baseline[3] = {0, 0, 0};
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Kmat[3 * r + c] = this->_camera_calibration_matrix(r, c);
  if (K == VSLAM_ALIGNER_STEREO_UV) {
    const Vector3 b = frame_current_->cameraRight()->baselineHomogeneous();   // S:66
    baseline[0] = b(0); baseline[1] = b(1); baseline[2] = b(2);
  }
  check(vslam_aligner_upload(_handle, n, moving, fixed, omega, wt, Kmat, baseline, (int32_t)this->_number_of_rows_image,
                             (int32_t)this->_number_of_cols_image, this->_minimum_reliable_depth_meters));
}

template <class A, int K>
void GpuFrameAligner<A, K>::fetch_errors_and_inliers() {
  const size_t n = this->_number_of_measurements;
  this->_errors.resize(n);
  _inlier_bytes.resize(n);
  check(vslam_aligner_download(_handle, this->_errors.data(), _inlier_bytes.data()));
  this->_inliers.resize(n);
  for (size_t u = 0; u < n; ++u) this->_inliers[u] = _inlier_bytes[u] != 0;   // std::vector<bool>
}

template <class A, int K>
void GpuFrameAligner<A, K>::linearize(const bool& ignore_outliers_) {        // S:72-187 / U:77-171
  double T[12];
  pose_to_array(T);
  vslam_linear_system s;
  check(vslam_aligner_linearize(_handle, T, ignore_outliers_, this->_parameters->maximum_error_kernel, &s));
  adopt(s);
  fetch_errors_and_inliers();
}

template <class A, int K>
void GpuFrameAligner<A, K>::oneRound(const bool& ignore_outliers_) {         // S:190-207 / U:174-191
  double T[12];
  pose_to_array(T);
  vslam_linear_system s;
  const vslam_aligner_parameters q = parameters_for_abi();
  check(vslam_aligner_one_round(_handle, &q, ignore_outliers_, T, &s));
  adopt(s);
  array_to_pose(T);
}

template <class A, int K>
void GpuFrameAligner<A, K>::converge() {                                      // S:210-264 / U:194-248
  double T[12], information[36];
  pose_to_array(T);
  vslam_linear_system s;
  int32_t converged = 0, rounds = 0;
  const vslam_aligner_parameters q = parameters_for_abi();
  // the whole Gauss-Newton loop as one persistent device kernel (bit-identical to the round-by-round driver)
  check(vslam_aligner_converge_fused(_handle, &q, T, &s, information, &converged, &rounds));
  adopt(s);
  array_to_pose(T);
  this->_has_system_converged = converged != 0;
  if (converged)
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c) this->_information_matrix(r, c) = information[6 * r + c];   // S:239
  fetch_errors_and_inliers();   // PoseTracker3D::_prunePoints reads errors() / inliers() (pose_tracker_3d.cpp:437-472)

  // VISUALIZATION ONLY (S:258-263): cheap, host side
  for (Index u = 0; u < this->_number_of_measurements; ++u) {
    FramePoint* frame_point = this->_frame_current->points()[u];
    ImageCoordinates image_coordinates(this->_camera_calibration_matrix * this->_previous_to_current * this->_moving[u]);
    image_coordinates /= image_coordinates.z();
    frame_point->setProjectionEstimateLeftOptimized(cv::Point2f(image_coordinates.x(), image_coordinates.y()));
  }
}

template class GpuFrameAligner<StereoUVAligner, VSLAM_ALIGNER_STEREO_UV>;
template class GpuFrameAligner<UVDAligner, VSLAM_ALIGNER_UVD>;

}  // namespace proslam
