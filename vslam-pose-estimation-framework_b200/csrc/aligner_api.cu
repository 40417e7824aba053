// aligner_api.cu -- C ABI of the frame aligners (include/vslam_b200.h): device-resident SoA correspondences, the
// linearize launch, and the host-side Gauss-Newton driver mirroring BaseAligner::oneRound / converge
// (reference src/aligners/stereouv_aligner.cpp:190-264, uvd_aligner.cpp:174-248).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/vslam_b200.h"
#include "api_common.h"
#include "gn_math.h"
#include "host_math.h"
#include "kernels.cuh"

using namespace vslam;

struct vslam_aligner {
  int kind = 0;
  int device = 0;
  int sm_count = 0;
  int max_points = 0;
  int n = 0;
  int fixed_dim = 4, omega_dim = 1;
  cudaStream_t stream = nullptr;
  double* d_moving = nullptr;
  double* d_fixed = nullptr;
  double* d_omega = nullptr;
  double* d_wt = nullptr;
  double* d_errors = nullptr;
  uint8_t* d_inliers = nullptr;
  double* d_partials = nullptr;
  double* d_system = nullptr;
  unsigned int* d_ticket = nullptr;
  double* h_stage = nullptr;    // pinned SoA staging, (3 + fixed_dim + omega_dim + 1) * max_points
  double* h_system = nullptr;   // pinned [32]
  AlignerCamera cam;
  int max_grid = 0;
  int resident_blocks = 0;      // co-resident CTAs of the cooperative kernel = grid cap of both linearize paths
  GnControl* d_ctl = nullptr;
  GnControl* h_ctl = nullptr;   // pinned
  int64_t launches = 0;
  bool uploaded = false;
};

namespace {

AlignerBuffers buffers(const vslam_aligner* h) {
  AlignerBuffers b;
  b.moving = h->d_moving;
  b.fixed = h->d_fixed;
  b.omega = h->d_omega;
  b.wt = h->d_wt;
  b.errors = h->d_errors;
  b.inliers = h->d_inliers;
  b.partials = h->d_partials;
  b.system = h->d_system;
  b.ticket = h->d_ticket;
  b.stride = h->max_points;
  return b;
}

void unpack_system(const vslam_aligner* h, vslam_linear_system* s) {
  const double* v = h->h_system;
  int k = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j, ++k) s->H[i * 6 + j] = s->H[j * 6 + i] = v[k];
  for (int i = 0; i < 6; ++i) s->b[i] = v[21 + i];
  s->total_error = v[27];
  s->number_of_inliers = (int32_t)std::llrint(v[28]);
  s->number_of_outliers = h->n - s->number_of_inliers;   // stereouv :186 / uvd :170
}

int linearize_async(vslam_aligner* h, const double T[12], int ignore_outliers, double kernel) {
  if (!h->uploaded) return fail(VSLAM_ERR_STATE, "linearize before upload");
  CUDA_TRY(cudaSetDevice(h->device));
  if (h->n == 0) {
    CUDA_TRY(cudaMemsetAsync(h->d_system, 0, sizeof(double) * 32, h->stream));
    return VSLAM_OK;
  }
  launch_linearize(h->kind, h->n, buffers(h), h->cam, T, ignore_outliers, kernel, aligner_grid(h->n, h->resident_blocks),
                   h->stream);
  ++h->launches;
  CUDA_TRY(cudaGetLastError());
  return VSLAM_OK;
}

int read_system(vslam_aligner* h, vslam_linear_system* s) {
  CUDA_TRY(cudaMemcpyAsync(h->h_system, h->d_system, sizeof(double) * 32, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (s) unpack_system(h, s);
  return VSLAM_OK;
}

int one_round(vslam_aligner* h, const vslam_aligner_parameters* p, int ignore_outliers, double T[12], vslam_linear_system* s) {
  int rc = linearize_async(h, T, ignore_outliers, p->maximum_error_kernel);   // :193 / :177
  if (rc) return rc;
  if ((rc = read_system(h, s))) return rc;
  for (int i = 0; i < 6; ++i) s->H[i * 6 + i] += p->damping * h->n;           // :196 / :180
  double nb[6], dx[6];
  for (int i = 0; i < 6; ++i) nb[i] = -s->b[i];
  solve6(s->H, nb, dx);                                                       // :199 / :183
  apply_update(dx, T);                                                        // :200-206 / :184-190
  return VSLAM_OK;
}

}  // namespace

extern "C" {

int vslam_aligner_create(int kind, int32_t max_points, int device, vslam_aligner** out) {
  if (!out) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  if (kind != VSLAM_ALIGNER_STEREO_UV && kind != VSLAM_ALIGNER_UVD) return fail(VSLAM_ERR_INVALID_ARGUMENT, "unknown aligner kind %d", kind);
  if (max_points < 1) return fail(VSLAM_ERR_INVALID_ARGUMENT, "max_points must be positive");
  int rc = require_device(device);
  if (rc) return rc;
  vslam_aligner* h = new vslam_aligner();
  h->kind = kind;
  h->device = device;
  h->max_points = (max_points + 31) & ~31;
  h->fixed_dim = kind == VSLAM_ALIGNER_STEREO_UV ? 4 : 3;
  h->omega_dim = kind == VSLAM_ALIGNER_STEREO_UV ? 1 : 2;
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
  h->resident_blocks = std::max(1, converge_max_blocks_per_sm(kind)) * h->sm_count;
  h->max_grid = aligner_grid(h->max_points, h->resident_blocks);
  const size_t N = h->max_points;
  bool ok = true;
  auto dalloc = [&](void** p, size_t bytes) {
    if (ok && cudaMalloc(p, bytes) != cudaSuccess) ok = false;
  };
  dalloc((void**)&h->d_moving, sizeof(double) * 3 * N);
  dalloc((void**)&h->d_fixed, sizeof(double) * h->fixed_dim * N);
  dalloc((void**)&h->d_omega, sizeof(double) * h->omega_dim * N);
  dalloc((void**)&h->d_wt, sizeof(double) * N);
  dalloc((void**)&h->d_errors, sizeof(double) * N);
  dalloc((void**)&h->d_inliers, N);
  dalloc((void**)&h->d_partials, sizeof(double) * 32 * h->max_grid);
  dalloc((void**)&h->d_system, sizeof(double) * 32);
  dalloc((void**)&h->d_ticket, sizeof(unsigned int));
  dalloc((void**)&h->d_ctl, sizeof(GnControl));
  if (ok && cudaMallocHost((void**)&h->h_ctl, sizeof(GnControl)) != cudaSuccess) ok = false;
  if (ok && cudaMallocHost((void**)&h->h_stage, sizeof(double) * (3 + h->fixed_dim + h->omega_dim + 1) * N) != cudaSuccess) ok = false;
  if (ok && cudaMallocHost((void**)&h->h_system, sizeof(double) * 32) != cudaSuccess) ok = false;
  if (ok && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) ok = false;
  if (ok && cudaMemset(h->d_ticket, 0, sizeof(unsigned int)) != cudaSuccess) ok = false;
  if (!ok) {
    const cudaError_t e = cudaGetLastError();
    vslam_aligner_destroy(h);
    return fail(VSLAM_ERR_CUDA, "device allocation failed: %s", cudaGetErrorString(e));
  }
  *out = h;
  return VSLAM_OK;
}

int vslam_aligner_destroy(vslam_aligner* h) {
  if (!h) return VSLAM_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  cudaFree(h->d_moving); cudaFree(h->d_fixed); cudaFree(h->d_omega); cudaFree(h->d_wt); cudaFree(h->d_errors);
  cudaFree(h->d_inliers); cudaFree(h->d_partials); cudaFree(h->d_system); cudaFree(h->d_ticket);
  cudaFree(h->d_ctl);
  cudaFreeHost(h->h_stage); cudaFreeHost(h->h_system); cudaFreeHost(h->h_ctl);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return VSLAM_OK;
}

int vslam_aligner_upload(vslam_aligner* h, int32_t n, const double* moving, const double* fixed, const double* omega,
                         const double* wt, const double K[9], const double baseline[3], int32_t rows, int32_t cols,
                         double minimum_depth) {
  if (!h) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null handle");
  if (n < 0 || n > h->max_points) return fail(VSLAM_ERR_CAPACITY, "n=%d outside [0, max_points=%d]", n, h->max_points);
  if (n && (!moving || !fixed || !omega || !wt)) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null correspondence array");
  if (!K) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null camera matrix");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));   // the staging buffer may still feed a previous copy
  const size_t N = h->max_points;
  // AoS (the reference's std::vector<Vector3>, ...) -> SoA planes, so that the kernel's loads coalesce
  double* s = h->h_stage;
  double* sm = s;
  double* sf = sm + 3 * N;
  double* so = sf + h->fixed_dim * N;
  double* sw = so + h->omega_dim * N;
  for (int u = 0; u < n; ++u) {
    for (int k = 0; k < 3; ++k) sm[k * N + u] = moving[3 * u + k];
    for (int k = 0; k < h->fixed_dim; ++k) sf[k * N + u] = fixed[h->fixed_dim * u + k];
    for (int k = 0; k < h->omega_dim; ++k) so[k * N + u] = omega[h->omega_dim * u + k];
    sw[u] = wt[u];
  }
  if (n) {
    CUDA_TRY(cudaMemcpyAsync(h->d_moving, sm, sizeof(double) * 3 * N, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_fixed, sf, sizeof(double) * h->fixed_dim * N, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_omega, so, sizeof(double) * h->omega_dim * N, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_wt, sw, sizeof(double) * N, cudaMemcpyHostToDevice, h->stream));
  }
  for (int i = 0; i < 9; ++i) h->cam.K[i] = K[i];
  for (int i = 0; i < 3; ++i) h->cam.baseline[i] = baseline ? baseline[i] : 0.0;
  h->cam.rows = rows;
  h->cam.cols = cols;
  h->cam.min_depth = minimum_depth;
  h->n = n;
  h->uploaded = true;
  return VSLAM_OK;
}

int vslam_aligner_linearize(vslam_aligner* h, const double T[12], int ignore_outliers, double kernel, vslam_linear_system* s) {
  if (!h || !T || !s) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  int rc = linearize_async(h, T, ignore_outliers, kernel);
  if (rc) return rc;
  return read_system(h, s);
}

int vslam_aligner_linearize_async(vslam_aligner* h, const double T[12], int ignore_outliers, double kernel) {
  if (!h || !T) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  return linearize_async(h, T, ignore_outliers, kernel);
}

int vslam_aligner_read_system(vslam_aligner* h, vslam_linear_system* s) {
  if (!h || !s) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  return read_system(h, s);
}

int vslam_aligner_download(vslam_aligner* h, double* errors, uint8_t* inliers) {
  if (!h) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  if (errors && h->n) CUDA_TRY(cudaMemcpyAsync(errors, h->d_errors, sizeof(double) * h->n, cudaMemcpyDeviceToHost, h->stream));
  if (inliers && h->n) CUDA_TRY(cudaMemcpyAsync(inliers, h->d_inliers, h->n, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return VSLAM_OK;
}

int vslam_aligner_one_round(vslam_aligner* h, const vslam_aligner_parameters* p, int ignore_outliers, double T[12],
                            vslam_linear_system* s) {
  if (!h || !p || !T || !s) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  return one_round(h, p, ignore_outliers, T, s);
}

int vslam_aligner_converge(vslam_aligner* h, const vslam_aligner_parameters* p, double T[12], vslam_linear_system* s,
                           double* information, int32_t* has_converged, int32_t* number_of_rounds) {
  if (!h || !p || !T || !s) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  double total_error_previous = 0;                                            // :213 / :197
  int converged = 0, rounds = 0, rc;
  // StereoUV gates the inlier-only phase on minimum_number_of_inliers (:224), UVD on the literal 100 (:208)
  const int inlier_gate = h->kind == VSLAM_ALIGNER_STEREO_UV ? p->minimum_number_of_inliers : 100;
  for (int it = 0; it < p->maximum_number_of_iterations; ++it) {              // :216 / :200
    if ((rc = one_round(h, p, 0, T, s))) return rc;
    ++rounds;
    if (p->error_delta_for_convergence > std::fabs(total_error_previous - s->total_error)) {
      total_error_previous = s->total_error;
      if (s->number_of_inliers > inlier_gate && s->number_of_inliers > s->number_of_outliers) {
        for (int ii = 0; ii < p->maximum_number_of_iterations; ++ii) {        // inlier-only rounds
          if ((rc = one_round(h, p, 1, T, s))) return rc;
          ++rounds;
          const bool done = std::fabs(total_error_previous - s->total_error) < p->error_delta_for_convergence;
          total_error_previous = s->total_error;
          if (done) break;
        }
      }
      if (information) std::memcpy(information, s->H, sizeof(double) * 36);   // _information_matrix = _H
      converged = 1;
      break;
    }
    total_error_previous = s->total_error;
  }
  if (has_converged) *has_converged = converged;
  if (number_of_rounds) *number_of_rounds = rounds;
  return VSLAM_OK;
}

int vslam_aligner_converge_fused(vslam_aligner* h, const vslam_aligner_parameters* p, double T[12], vslam_linear_system* s,
                                 double* information, int32_t* has_converged, int32_t* number_of_rounds) {
  if (!h || !p || !T || !s) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  if (!h->uploaded) return fail(VSLAM_ERR_STATE, "converge before upload");
  CUDA_TRY(cudaSetDevice(h->device));
  if (h->n == 0 || p->maximum_number_of_iterations < 1)   // nothing to iterate on the device: the stepwise driver
    return vslam_aligner_converge(h, p, T, s, information, has_converged, number_of_rounds);
  GnControl* c = h->h_ctl;
  std::memset(c, 0, sizeof(*c));
  for (int i = 0; i < 12; ++i) c->T[i] = T[i];
  GnParams gp;
  gp.error_delta = p->error_delta_for_convergence;
  gp.kernel = p->maximum_error_kernel;
  gp.damping = p->damping;
  gp.max_iterations = p->maximum_number_of_iterations;
  gp.inlier_gate = h->kind == VSLAM_ALIGNER_STEREO_UV ? p->minimum_number_of_inliers : 100;   // :224 / :208
  CUDA_TRY(cudaMemcpyAsync(h->d_ctl, c, sizeof(GnControl), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(launch_converge(h->kind, h->n, buffers(h), h->cam, gp, h->d_ctl, aligner_grid(h->n, h->resident_blocks), h->stream));
  ++h->launches;
  CUDA_TRY(cudaMemcpyAsync(c, h->d_ctl, sizeof(GnControl), cudaMemcpyDeviceToHost, h->stream));
  int rc = read_system(h, s);   // synchronises
  if (rc) return rc;
  for (int i = 0; i < 12; ++i) T[i] = c->T[i];
  for (int i = 0; i < 6; ++i) s->H[i * 6 + i] += p->damping * h->n;   // _H after oneRound carries the damping (:196)
  if (information && c->converged) std::memcpy(information, c->H, sizeof(double) * 36);
  if (has_converged) *has_converged = c->converged;
  if (number_of_rounds) *number_of_rounds = c->rounds;
  return VSLAM_OK;
}

void* vslam_aligner_stream(vslam_aligner* h) { return h ? (void*)h->stream : nullptr; }

int vslam_aligner_synchronize(vslam_aligner* h) {
  if (!h) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return VSLAM_OK;
}

int64_t vslam_aligner_launch_count(const vslam_aligner* h) { return h ? h->launches : 0; }

void vslam_solve6(const double A[36], const double rhs[6], double x[6]) { solve6(A, rhs, x); }
void vslam_v2t(const double v[6], double T[12]) { v2t(v, T); }

}  // extern "C"
