// aligner_api.cu -- C ABI of the frame aligners (include/vslam_b200.h): device-resident SoA correspondences, the
// linearize launch, and the host-side Gauss-Newton driver mirroring BaseAligner::oneRound / converge
// (reference src/aligners/stereouv_aligner.cpp:190-264, uvd_aligner.cpp:174-248).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/vslam_b200.h"
#include "aligner_internal.h"
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
  // ONE device block and its pinned mirror, laid out so that a frame costs one H2D copy, one kernel and one D2H copy:
  //   [ errors (stride f64) | inliers (stride u8, padded) ]  <- ends at kHeader; the OUT region of the last upload
  //   [ GnControl (512 B) | system (32 f64) ]                <- header at a fixed offset
  //   [ moving 3 | fixed 4|3 | omega 1|2 | wt 1 ] x stride   <- the IN region: SoA planes with stride = n rounded to 32
  // H2D = header + IN (contiguous), D2H = OUT + header (contiguous).  `stride` follows the problem size: a frame with
  // 700 tracks moves 50 KB, not planes of max_points entries.
  uint8_t* d_io = nullptr;
  uint8_t* h_io = nullptr;      // pinned
  size_t header_offset = 0;     // = out_bytes(max_points)
  int stride = 32;
  bool in_dirty = false;        // the pinned IN region is newer than the device's
  bool host_out_valid = false;  // the pinned OUT region holds errors / inliers of the last device run
  double* d_partials = nullptr;
  unsigned int* d_ticket = nullptr;
  AlignerCamera cam;
  int max_grid = 0;
  int resident_blocks = 0;      // co-resident CTAs of the cooperative kernel = grid cap of both linearize paths
  int64_t launches = 0;
  bool uploaded = false;
  double last_total_error = 0;
};

namespace {
constexpr size_t kCtlBytes = 512, kHeaderBytes = 768;
static_assert(sizeof(GnControl) <= kCtlBytes, "GnControl fits its slot");
size_t out_bytes(int stride) { return (size_t)stride * 8 + (((size_t)stride + 7) & ~(size_t)7); }
size_t in_bytes(const vslam_aligner* h, int stride) { return (size_t)stride * 8 * (3 + h->fixed_dim + h->omega_dim + 1); }
template <class T> T* at(uint8_t* base, size_t off) { return reinterpret_cast<T*>(base + off); }
GnControl* ctl_of(uint8_t* base, const vslam_aligner* h) { return at<GnControl>(base, h->header_offset); }
double* system_of(uint8_t* base, const vslam_aligner* h) { return at<double>(base, h->header_offset + kCtlBytes); }
double* errors_of(uint8_t* base, const vslam_aligner* h) { return at<double>(base, h->header_offset - out_bytes(h->stride)); }
uint8_t* inliers_of(uint8_t* base, const vslam_aligner* h) {
  return base + h->header_offset - out_bytes(h->stride) + (size_t)h->stride * 8;
}
double* planes_of(uint8_t* base, const vslam_aligner* h) { return at<double>(base, h->header_offset + kHeaderBytes); }
}  // namespace

namespace {

AlignerBuffers buffers(const vslam_aligner* h) {
  AlignerBuffers b;
  const size_t S = h->stride;
  double* planes = planes_of(h->d_io, h);
  b.moving = planes;
  b.fixed = planes + 3 * S;
  b.omega = planes + (3 + h->fixed_dim) * S;
  b.wt = planes + (3 + h->fixed_dim + h->omega_dim) * S;
  b.errors = errors_of(h->d_io, h);
  b.inliers = inliers_of(h->d_io, h);
  b.partials = h->d_partials;
  b.system = system_of(h->d_io, h);
  b.ticket = h->d_ticket;
  b.stride = h->stride;
  return b;
}

// header (+ correspondences when they changed) host -> device: ONE copy
int push_inputs(vslam_aligner* h, bool with_header) {
  if (h->in_dirty) {
    CUDA_TRY(cudaMemcpyAsync(h->d_io + h->header_offset, h->h_io + h->header_offset, kHeaderBytes + in_bytes(h, h->stride),
                             cudaMemcpyHostToDevice, h->stream));
    h->in_dirty = false;
  } else if (with_header) {
    CUDA_TRY(cudaMemcpyAsync(h->d_io + h->header_offset, h->h_io + h->header_offset, kCtlBytes, cudaMemcpyHostToDevice, h->stream));
  }
  return VSLAM_OK;
}

void unpack_system(const vslam_aligner* h, vslam_linear_system* s) {
  const double* v = system_of(h->h_io, h);
  int k = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j, ++k) s->H[i * 6 + j] = s->H[j * 6 + i] = v[k];
  for (int i = 0; i < 6; ++i) s->b[i] = v[21 + i];
  s->total_error = v[27];
  const_cast<vslam_aligner*>(h)->last_total_error = v[27];
  s->number_of_inliers = (int32_t)std::llrint(v[28]);
  s->number_of_outliers = h->n - s->number_of_inliers;   // stereouv :186 / uvd :170
}

int linearize_async(vslam_aligner* h, const double T[12], int ignore_outliers, double kernel) {
  if (!h->uploaded) return fail(VSLAM_ERR_STATE, "linearize before upload");
  CUDA_TRY(cudaSetDevice(h->device));
  h->host_out_valid = false;
  if (h->n == 0) {
    CUDA_TRY(cudaMemsetAsync(system_of(h->d_io, h), 0, sizeof(double) * 32, h->stream));
    return VSLAM_OK;
  }
  int rc = push_inputs(h, false);
  if (rc) return rc;
  launch_linearize(h->kind, h->n, buffers(h), h->cam, T, ignore_outliers, kernel, aligner_grid(h->n, h->resident_blocks),
                   h->stream);
  ++h->launches;
  CUDA_TRY(cudaGetLastError());
  return VSLAM_OK;
}

int read_system(vslam_aligner* h, vslam_linear_system* s) {
  CUDA_TRY(cudaMemcpyAsync(system_of(h->h_io, h), system_of(h->d_io, h), sizeof(double) * 32, cudaMemcpyDeviceToHost, h->stream));
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

namespace vslam {
int fetch_aligner_results(vslam_aligner* h, AlignerResults* out) {
  if (!h || !out) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  if (!h->uploaded) return fail(VSLAM_ERR_STATE, "the aligner holds no correspondences");
  CUDA_TRY(cudaSetDevice(h->device));
  if (h->n && !h->host_out_valid) {
    CUDA_TRY(cudaMemcpyAsync(errors_of(h->h_io, h), errors_of(h->d_io, h), out_bytes(h->stride), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->host_out_valid = true;
  }
  out->n = h->n;
  out->device = h->device;
  out->d_errors = errors_of(h->d_io, h);
  out->d_inliers = inliers_of(h->d_io, h);
  out->h_errors = errors_of(h->h_io, h);
  out->h_inliers = inliers_of(h->h_io, h);
  out->total_error = h->last_total_error;
  return VSLAM_OK;
}
}  // namespace vslam

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
  h->header_offset = (out_bytes((int)N) + 255) & ~(size_t)255;
  const size_t io_bytes = h->header_offset + kHeaderBytes + in_bytes(h, (int)N);
  dalloc((void**)&h->d_io, io_bytes);
  dalloc((void**)&h->d_partials, sizeof(double) * 32 * h->max_grid);
  dalloc((void**)&h->d_ticket, sizeof(unsigned int));
  if (ok && cudaMallocHost((void**)&h->h_io, io_bytes) != cudaSuccess) ok = false;
  if (ok) std::memset(h->h_io + h->header_offset, 0, kHeaderBytes);
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
  cudaFree(h->d_io); cudaFree(h->d_partials); cudaFree(h->d_ticket);
  cudaFreeHost(h->h_io);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return VSLAM_OK;
}

int vslam_aligner_upload(vslam_aligner* h, int32_t n, const double* moving, const double* fixed, const double* omega,
                         const double* wt, const double K[9], const double baseline[3], int32_t rows, int32_t cols,
                         double minimum_depth) {
  VSLAM_NVTX("vslam_aligner_upload [StereoUVAligner::initialize]");
  if (!h) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null handle");
  if (n < 0 || n > h->max_points) return fail(VSLAM_ERR_CAPACITY, "n=%d outside [0, max_points=%d]", n, h->max_points);
  if (n && (!moving || !fixed || !omega || !wt)) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null correspondence array");
  if (!K) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null camera matrix");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));   // the staging buffer may still feed a previous copy
  // AoS (the reference's std::vector<Vector3>, ...) -> SoA planes with stride = n rounded up to 32, so that the kernel's
  // loads coalesce and the copy moves the problem, not the capacity.  The copy itself travels with the next launch.
  h->stride = std::max(32, (n + 31) & ~31);
  const size_t N = h->stride;
  double* sm = planes_of(h->h_io, h);
  double* sf = sm + 3 * N;
  double* so = sf + h->fixed_dim * N;
  double* sw = so + h->omega_dim * N;
  for (int u = 0; u < n; ++u) {
    for (int k = 0; k < 3; ++k) sm[k * N + u] = moving[3 * u + k];
    for (int k = 0; k < h->fixed_dim; ++k) sf[k * N + u] = fixed[h->fixed_dim * u + k];
    for (int k = 0; k < h->omega_dim; ++k) so[k * N + u] = omega[h->omega_dim * u + k];
    sw[u] = wt[u];
  }
  h->in_dirty = n > 0;
  h->host_out_valid = false;
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
  VSLAM_NVTX("vslam_aligner_linearize");
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
  VSLAM_NVTX("vslam_aligner_download [errors / inliers]");
  if (!h) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  if (h->n == 0) return VSLAM_OK;
  if (!h->host_out_valid) {   // (the fused converge brings errors / inliers back with its own result copy)
    CUDA_TRY(cudaMemcpyAsync(errors_of(h->h_io, h), errors_of(h->d_io, h), out_bytes(h->stride), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->host_out_valid = true;
  }
  if (errors) std::memcpy(errors, errors_of(h->h_io, h), sizeof(double) * h->n);
  if (inliers) std::memcpy(inliers, inliers_of(h->h_io, h), h->n);
  return VSLAM_OK;
}

int vslam_aligner_one_round(vslam_aligner* h, const vslam_aligner_parameters* p, int ignore_outliers, double T[12],
                            vslam_linear_system* s) {
  VSLAM_NVTX("vslam_aligner_one_round");
  if (!h || !p || !T || !s) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  return one_round(h, p, ignore_outliers, T, s);
}

int vslam_aligner_converge(vslam_aligner* h, const vslam_aligner_parameters* p, double T[12], vslam_linear_system* s,
                           double* information, int32_t* has_converged, int32_t* number_of_rounds) {
  VSLAM_NVTX("vslam_aligner_converge [PoseTracker3D::compute->pose_optim]");
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
  VSLAM_NVTX("vslam_aligner_converge_fused [PoseTracker3D::compute->pose_optim]");
  if (!h || !p || !T || !s) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  if (!h->uploaded) return fail(VSLAM_ERR_STATE, "converge before upload");
  CUDA_TRY(cudaSetDevice(h->device));
  if (h->n == 0 || p->maximum_number_of_iterations < 1)   // nothing to iterate on the device: the stepwise driver
    return vslam_aligner_converge(h, p, T, s, information, has_converged, number_of_rounds);
  GnControl* c = ctl_of(h->h_io, h);
  std::memset(c, 0, sizeof(*c));
  for (int i = 0; i < 12; ++i) c->T[i] = T[i];
  GnParams gp;
  gp.error_delta = p->error_delta_for_convergence;
  gp.kernel = p->maximum_error_kernel;
  gp.damping = p->damping;
  gp.max_iterations = p->maximum_number_of_iterations;
  gp.inlier_gate = h->kind == VSLAM_ALIGNER_STEREO_UV ? p->minimum_number_of_inliers : 100;   // :224 / :208
  // one copy in (control block + the correspondences if they are new), one kernel, one copy out (errors, inliers, control
  // block, system) and one synchronisation
  int rc = push_inputs(h, true);
  if (rc) return rc;
  CUDA_TRY(launch_converge(h->kind, h->n, buffers(h), h->cam, gp, ctl_of(h->d_io, h), aligner_grid(h->n, h->resident_blocks), h->stream));
  ++h->launches;
  CUDA_TRY(cudaMemcpyAsync(errors_of(h->h_io, h), errors_of(h->d_io, h), out_bytes(h->stride) + kHeaderBytes,
                           cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->host_out_valid = true;
  unpack_system(h, s);
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
