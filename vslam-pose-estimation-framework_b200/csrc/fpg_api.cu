// fpg_api.cu -- C ABI of the stereo framepoint generator (include/vslam_b200.h): handle lifetime, device buffers,
// streams, chunked batch pipeline with copy/compute overlap, and the host-side scalar logic of
// StereoFramePointGenerator::{configure, initialize, compute} (reference
// src/framepoint_generation/stereo_framepoint_generator.cpp:16-60, 73-133, 135-462).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/vslam_b200.h"
#include "aligner_internal.h"
#include "api_common.h"
#include "host_math.h"
#include "kernels.cuh"

using namespace vslam;

static_assert(sizeof(FramePointRecord) == sizeof(vslam_framepoint), "record layout");
static_assert(sizeof(TrackedPoint) == sizeof(vslam_tracked_point), "tracked layout");
static_assert(sizeof(PreviousPoint) == sizeof(vslam_previous_point) && sizeof(PreviousPoint) == 128, "previous point layout");
static_assert(sizeof(TrackRecord) == sizeof(vslam_track) && sizeof(TrackRecord) == 88, "track layout");
static_assert(sizeof(RecoveredRecord) == sizeof(vslam_recovered_point) && sizeof(RecoveredRecord) == 112, "recovered layout");

namespace {

constexpr int kLanes = 2;   // chunk pipelines that overlap copies and the small kernels of neighbouring chunks

struct Lane {
  cudaStream_t stream = nullptr;

  uint8_t* blurred = nullptr;   // [2*chunk][rows][pitch]
  uint32_t* mask = nullptr;     // [2*chunk][rows][mask_words]
  CUtensorMap blurred_map;      // TMA descriptor of `blurred` (box = describe tile)
  uint8_t* stage = nullptr;     // [2][chunk * pair_stride] dense landing zone of the H2D copies
  size_t stage_bytes = 0;       // per side
};

// profiling: events at the kernel boundaries of one chunk (single lane, serialised)
enum { kEvFast0, kEvFast1, kEvCompact1, kEvBlur1, kEvDescribe1, kEvMatch0, kEvMatch1, kEvSelect1, kEvLin0, kEvLin1, kEvTrack0, kEvTrack1, kNumEv };
enum { kKFast, kKCompact, kKBlur, kKDescribe, kKMatch, kKSelect, kKLinearize, kKTrack, kKFrameAligner, kNumKernels };
static_assert(kNumKernels == VSLAM_FPG_KERNELS, "kernel profile size");

struct StageClock {
  cudaEvent_t ev[kNumEv] = {};
};

}  // namespace

struct vslam_fpg {
  vslam_fpg_config cfg;
  int device = 0;
  int sm_count = 0;
  Geometry g;
  int n_passes = 1;
  int chunk = 1;              // pairs per pipeline chunk
  int describe_sub_chunk = 64;   // pairs per blur -> rBRIEF launch pair (L2 residency of the blurred scratch)
  int max_batch = 1;
  int out_cap = 0;            // framepoint records per pair
  HostRegion regions[kMaxRegions];
  double thresholds[kMaxRegions];
  int target_keypoints = 0;
  int target_per_detector = 0;
  StereoParams sp;
  Lane lanes[kLanes];
  Buffers b;                  // b.blurred / b.mask are per-lane and set per launch
  CUtensorMap image_map;      // TMA descriptor of b.image with the FAST tile as box
  CUtensorMap blur_map;       // ... with the blur tile as box
  FramePointRecord* d_out = nullptr;       // [max_batch][out_cap]
  FramePointRecord* d_matches = nullptr;   // [cap] (single pair, emission order)
  int32_t* d_n_matches = nullptr;
  TrackedPoint* d_tracked = nullptr;
  int tracked_cap = 0;
  // pinned staging
  int32_t* h_counts = nullptr;   // [2*max_batch][n_regions]
  int32_t* h_n_desc = nullptr;   // [2*max_batch]
  int32_t* h_n_out = nullptr;    // [max_batch][2]
  FramePointRecord* h_out_stage = nullptr;   // pinned, [out_cap]: the records of the single-pair compute(), copied with the counts
  int32_t* h_flag = nullptr;
  int32_t* d_status = nullptr;   // [error_flag, pad | n_desc 2B | raw_count 2B x regions]; b.error_flag / n_desc / raw_count point into it
  int32_t* h_status = nullptr;   // pinned mirror; h_flag / h_n_desc / h_counts point into it
  int32_t* d_step_bins = nullptr;       // [rows_bin * cols_bin + 1] winners of the match replay (fused frame)
  LandmarkEstimate* d_step_estimates = nullptr;   // [step_cap] landmark estimates of the device-resident points()
  int32_t step_points = 0;              // points() the device holds for the next fused frame
  int32_t* h_status_device = nullptr;   // the same words as the device addresses them (written by the fused frame)
  // state of the last single-pair initialize / last batch
  bool initialized = false;
  int last_pairs = 0;
  // features of the single-pair initialize(), downloaded behind its back on lane 1's stream (prefetch_features):
  // packed by pack_features_kernel (layout: feature_pack_bytes in kernels.cuh), device buffer + pinned host copy
  // The device side of the single-pair initialize() -- threshold upload, repitch, FAST, compact, blur / box sums,
  // descriptors, status download: ~10 stream operations of a few microseconds each -- is captured ONCE into a CUDA graph
  // and re-launched per frame.  Nothing in it changes from frame to frame: the thresholds travel through a pinned
  // buffer that a copy node of the graph reads, the images are copied into the staging buffer before the launch.
  cudaGraphExec_t init_graph = nullptr;
  const uint8_t* init_graph_stage = nullptr;   // the graph is valid for this staging buffer ...
  size_t init_graph_stride = 0;                // ... and this row stride
  int init_graph_kernels = 0;
  int64_t graph_launches = 0;
  bool init_graph_ok = true;                   // false after a failed capture: the direct launches are used
  bool init_graph_profiling = false;           // the graph was captured with / without the timing events
  bool capturing = false;
  int32_t* h_thr = nullptr;                    // pinned [kMaxRegions]
  int32_t* d_thr = nullptr;
  uint8_t* h_feat = nullptr;
  uint8_t* d_feat = nullptr;
  cudaEvent_t feat_ev = nullptr;
  bool feat_valid = false;
  int localizing = 1;
  double matching_distance = 0;
  int64_t launches = 0;
  bool profiling = false;
  StageClock clock;
  double t_detect = 0, t_describe = 0, t_match = 0;
  double k_ms[kNumKernels] = {};
  int64_t k_n[kNumKernels] = {};
  cudaEvent_t fork_ev = nullptr;
  cudaEvent_t join_ev[kLanes] = {};
  // single frames leave most of the 148 SMs idle: kernels that do not depend on each other (blur beside FAST + compact;
  // the epipolar match beside the aligner; the tracks' share of points() beside the bin selection) run on a side stream,
  // i.e. as parallel branches of the captured frame graph
  cudaStream_t side_stream = nullptr;
  cudaEvent_t branch_ev[4] = {};
  bool branches = true;                        // VSLAM_NO_FRAME_BRANCHES=1: one chain (A/B measurements)
  // batched StereoUV linearize over the pairs' own framepoints
  double* d_systems = nullptr;        // [max_batch][32]
  double* d_pair_errors = nullptr;    // [max_batch][out_cap]
  uint8_t* d_pair_inliers = nullptr;  // [max_batch][out_cap]
  double* h_systems = nullptr;        // pinned [max_batch][32]
  // track() / recoverPoints() (single pair)
  PreviousPoint* d_previous = nullptr;   // [previous_cap]
  int previous_cap = 0;
  TrackScratch track_scratch = {};
  TrackRecord* d_tracks = nullptr;       // [previous_cap]
  int32_t* d_lost = nullptr;             // [previous_cap]
  int32_t* h_track_stats = nullptr;      // pinned [4]
  TrackRecord* h_tracks = nullptr;       // pinned [previous_cap]
  int32_t* h_lost = nullptr;             // pinned [previous_cap]
  int n_device_tracks = -1;              // tracks of the last vslam_fpg_track still valid for compute()
  uint32_t* d_recover_xy = nullptr;      // [2][previous_cap]
  uint8_t* d_recover_desc = nullptr;     // [2][previous_cap][32] + [previous_cap] flags
  RecoveredRecord* d_recovered = nullptr;
  int32_t* d_recover_n = nullptr;        // {n_xy left, n_xy right, n_recovered}
  int8_t* d_brief_tests = nullptr;       // [256][4] when descriptor_type == VSLAM_DESCRIPTOR_BRIEF
  // fused tracked frame (vslam_fpg_frame_step): device-resident counters, the StereoUV aligner's planes, the pinned and
  // device-mapped result block, and one captured graph per frame status
  int step_cap = 0;                      // points per frame = cluster blocks x 256; 0 = not set up, -1 = unavailable
  int step_cluster_blocks = 0;
  FrameStepState* d_step = nullptr;
  GnControl* d_step_ctl = nullptr;
  double* d_step_planes = nullptr;       // [3 + 4 + 1 + 1][step_cap] moving | fixed | omega | wt
  double* d_step_errors = nullptr;       // [step_cap]
  uint8_t* d_step_inliers = nullptr;     // [step_cap]
  double* d_step_system = nullptr;       // [32]
  int32_t* d_step_track_length = nullptr;
  int32_t* d_step_kept_pos = nullptr;
  uint8_t* h_step = nullptr;             // pinned + mapped result block
  uint8_t* h_step_device = nullptr;      // its device address
  double* h_step_T = nullptr;            // pinned: the copied prefix of FrameStepState (motion prior, frame number, ticket = 0,
                                         // thresholds) that ONE copy node of the graph reads
  int32_t step_frame_id = 0;
  bool step_poll = true;                 // VSLAM_FRAME_STEP_SYNC=1: wait with cudaStreamSynchronize instead of polling
  size_t step_off_tracks = 0, step_off_kept = 0, step_off_errors = 0, step_off_inliers = 0, step_off_lost = 0,
         step_off_points = 0, step_off_frame_points = 0;
  cudaGraphExec_t step_graph[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [localizing][inbox buffer]
  int step_graph_kernels[2][2] = {{0, 0}, {0, 0}};
  // The images of a fused frame land in one of TWO inbox buffers (left at [0], right at [inbox_bytes]); the graph of a
  // frame reads the buffer its images are in.  vslam_fpg_frame_step_prefetch uploads the NEXT frame into the other one on
  // a copy stream of its own, i.e. while the current frame runs.
  uint8_t* step_inbox[2] = {nullptr, nullptr};
  size_t step_inbox_bytes = 0;           // per image
  size_t step_inbox_stride[2] = {0, 0};  // row stride of the staged pair
  int step_inbox_head = 0, step_inbox_count = 0;   // staged pairs: buffers head, head + 1 (mod 2)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t step_inbox_ev[2] = {nullptr, nullptr};
  const uint8_t* step_graph_stage = nullptr;
  size_t step_graph_stride = 0;
  bool step_graph_profiling = false;
  vslam_frame_step_parameters step_graph_parameters;
};

namespace {

// the captured frame graphs bake buffer addresses: any reallocation of a buffer they touch drops them
void invalidate_step_graphs(vslam_fpg* h) {
  for (auto& row : h->step_graph)
  for (auto& gph : row)
    if (gph) {
      cudaGraphExecDestroy(gph);
      gph = nullptr;
    }
}

void refresh_region_table(const vslam_fpg* h, RegionTable* rt) {
  for (int i = 0; i < h->g.n_regions; ++i) {
    rt->r[i] = Region{h->regions[i].x, h->regions[i].y, h->regions[i].w, h->regions[i].h};
    int t = (int)std::rint(h->thresholds[i]);   // FastDetector::setThreshold -> std::rint (:21); cv::FAST clamps
    rt->threshold[i] = std::min(std::max(t, 0), 255);
  }
}

int check_flag(vslam_fpg* h) {
  if (*h->h_flag == 1)
    return fail(VSLAM_ERR_CAPACITY, "more than max_keypoints_per_image=%d descriptor-valid keypoints in an image", h->g.cap);
  if (*h->h_flag == 2) return fail(VSLAM_ERR_CAPACITY, "framepoint output capacity exceeded");
  return VSLAM_OK;
}

inline void mark(vslam_fpg* h, Lane& lane, int ev) {
  if (!h->profiling) return;
  // inside a stream capture the timing events become external event-record nodes of the graph
  if (h->capturing) cudaEventRecordWithFlags(h->clock.ev[ev], lane.stream, cudaEventRecordExternal);
  else cudaEventRecord(h->clock.ev[ev], lane.stream);
}

// `to` continues after everything issued on `from` so far (inside a stream capture: an edge of the graph)
inline void order_after(vslam_fpg* h, cudaStream_t from, cudaStream_t to, int ev) {
  cudaEventRecord(h->branch_ev[ev], from);
  cudaStreamWaitEvent(to, h->branch_ev[ev], 0);
}

// single frames on lane 0 without the timing events (they bracket the kernels of ONE chain)
inline bool use_branches(const vslam_fpg* h, const Lane& lane, int n) {
  return h->branches && n == 1 && !h->profiling && &lane == &h->lanes[0];
}

// kernels of initialize() for images [2*p0, 2*(p0+n)) on one lane
void run_detect_describe(vslam_fpg* h, Lane& lane, int p0, int n, const int32_t* device_thresholds = nullptr,
                         bool counts_cleared = false, bool mask_cleared = false) {
  Buffers b = h->b;
  // lane-local scratch is indexed from image 0 of the chunk: shift the base so that image index 2*p0 lands on it
  b.blurred = lane.blurred - (size_t)2 * p0 * h->g.rows * h->g.pitch;
  b.mask = lane.mask - (size_t)2 * p0 * h->g.rows * h->g.mask_words;
  RegionTable rt;
  refresh_region_table(h, &rt);
  // (the pruned / consumed flags of the new frame are cleared by compact_kernel)
  const bool blur_beside_fast = use_branches(h, lane, n) && !h->d_brief_tests;
  if (blur_beside_fast) {   // the blur reads the image only: it does not wait for FAST + compact
    order_after(h, lane.stream, h->side_stream, 0);
    launch_blur(h->g, b, h->blur_map, 2 * p0, 2, h->side_stream);
  }
  mark(h, lane, kEvFast0);
  launch_fast(h->g, rt, b, h->image_map, 2 * p0, 2 * n, lane.stream, device_thresholds, counts_cleared, mask_cleared);
  mark(h, lane, kEvFast1);
  launch_compact(h->g, b, 2 * p0, 2 * n, lane.stream);
  mark(h, lane, kEvCompact1);
  if (blur_beside_fast) {
    order_after(h, h->side_stream, lane.stream, 0);
    launch_describe(h->g, b, lane.blurred_map, 2 * p0, 2, lane.stream, 0);
  } else if (h->d_brief_tests) {   // BRIEF-32: 9x9 box sums (u16) take the place of the blurred image
    uint16_t* boxsum = reinterpret_cast<uint16_t*>(lane.blurred);
    launch_box9(h->g, h->b.image + (size_t)2 * p0 * h->g.rows * h->g.pitch, boxsum, 2 * n, lane.stream);
    mark(h, lane, kEvBlur1);
    launch_describe_brief(h->g, boxsum, h->d_brief_tests, h->b.kp_xy + (size_t)2 * p0 * h->g.cap, h->b.n_desc + 2 * p0,
                          h->b.desc + (size_t)2 * p0 * h->g.cap * kDescBytes, h->g.cap, 2 * n, lane.stream);
  } else {
    // blur -> rBRIEF in sub-chunks that all use the FIRST images of the lane's blurred scratch: 64 pairs are 61 MB, which
    // stay in the 126 MB L2 between the two kernels and are overwritten there by the next sub-chunk, so the blurred
    // pixels (246 MB per 256-pair chunk, 100 % non-algorithmic) mostly never travel to DRAM.  (The timing events of the
    // profiling mode want the two kernels back to back: one sub-chunk then.)
    const int sub = h->profiling || h->capturing ? n : std::max(1, std::min(n, h->describe_sub_chunk));
    for (int q0 = 0; q0 < n; q0 += sub) {
      const int qn = std::min(sub, n - q0);
      Buffers bq = b;
      bq.blurred = lane.blurred - (size_t)2 * (p0 + q0) * h->g.rows * h->g.pitch;   // image 2 (p0 + q0) -> scratch image 0
      launch_blur(h->g, bq, h->blur_map, 2 * (p0 + q0), 2 * qn, lane.stream);
      if (q0 == 0) mark(h, lane, kEvBlur1);
      launch_describe(h->g, bq, lane.blurred_map, 2 * (p0 + q0), 2 * qn, lane.stream, 0);
      h->launches += q0 ? 2 : 0;
    }
  }
  mark(h, lane, kEvDescribe1);
  h->launches += 4;
}

// kernels of compute() for pairs [p0, p0+n) on one lane
void run_match_select(vslam_fpg* h, Lane& lane, int p0, int n, const TrackedPoint* tracked, int n_tracked,
                      bool generic_select = false) {
  mark(h, lane, kEvMatch0);
  for (int pass = 0; pass < h->n_passes; ++pass) {
    const int offset = pass == 0 ? 0 : ((pass & 1) ? (pass + 1) / 2 : -(pass / 2));
    launch_match(h->g, h->sp, h->b, p0, n, pass, offset, lane.stream);
    ++h->launches;
  }
  mark(h, lane, kEvMatch1);
  launch_select(h->g, h->sp, h->b, p0, n, h->n_passes, tracked, n_tracked, h->d_out, h->out_cap, generic_select,
                lane.stream);
  ++h->launches;
  mark(h, lane, kEvSelect1);
}

void add_interval(vslam_fpg* h, int kernel, int e0, int e1, int launches) {
  float ms = 0;
  if (cudaEventElapsedTime(&ms, h->clock.ev[e0], h->clock.ev[e1]) == cudaSuccess) {
    h->k_ms[kernel] += ms;
    h->k_n[kernel] += launches;
  } else {
    cudaGetLastError();
  }
}

void collect_clock(vslam_fpg* h, bool detect, bool match) {
  if (!h->profiling) return;
  if (detect) {
    const double d0 = h->k_ms[kKFast] + h->k_ms[kKCompact], e0 = h->k_ms[kKBlur] + h->k_ms[kKDescribe];
    add_interval(h, kKFast, kEvFast0, kEvFast1, 1);
    add_interval(h, kKCompact, kEvFast1, kEvCompact1, 1);
    add_interval(h, kKBlur, kEvCompact1, kEvBlur1, 1);
    add_interval(h, kKDescribe, kEvBlur1, kEvDescribe1, 1);
    h->t_detect += (h->k_ms[kKFast] + h->k_ms[kKCompact] - d0) * 1e-3;
    h->t_describe += (h->k_ms[kKBlur] + h->k_ms[kKDescribe] - e0) * 1e-3;
  }
  if (match) {
    const double m0 = h->k_ms[kKMatch] + h->k_ms[kKSelect];
    add_interval(h, kKMatch, kEvMatch0, kEvMatch1, h->n_passes);
    add_interval(h, kKSelect, kEvMatch1, kEvSelect1, 1);
    h->t_match += (h->k_ms[kKMatch] + h->k_ms[kKSelect] - m0) * 1e-3;
  }
}

// lanes other than 0 start after everything already ordered on lane 0 (the handle's stream) ...
int fork_lanes(vslam_fpg* h) {
  // an outstanding single-pair feature prefetch on lane 1 still reads the buffers a batched call overwrites
  if (h->feat_valid) CUDA_TRY(cudaStreamWaitEvent(h->lanes[0].stream, h->feat_ev, 0));
  CUDA_TRY(cudaEventRecord(h->fork_ev, h->lanes[0].stream));
  for (int l = 1; l < kLanes; ++l) CUDA_TRY(cudaStreamWaitEvent(h->lanes[l].stream, h->fork_ev, 0));
  return VSLAM_OK;
}

// ... and lane 0 continues only after they are done: work issued by a batched call is ordered on lane 0
int join_lanes(vslam_fpg* h) {
  for (int l = 1; l < kLanes; ++l) {
    CUDA_TRY(cudaEventRecord(h->join_ev[l], h->lanes[l].stream));
    CUDA_TRY(cudaStreamWaitEvent(h->lanes[0].stream, h->join_ev[l], 0));
  }
  return VSLAM_OK;
}

bool linear_upload(const Geometry& g, int n, size_t stride, size_t pair_stride) {
  return pair_stride == stride * (size_t)g.rows && (n > 1 || stride <= (size_t)g.cols + (size_t)g.cols / 4);
}

int ensure_stage(vslam_fpg* h, Lane& lane, size_t bytes, size_t pair_stride) {
  if (lane.stage_bytes >= bytes) return VSLAM_OK;
  CUDA_TRY(cudaStreamSynchronize(lane.stream));
  cudaFree(lane.stage);
  lane.stage = nullptr;
  lane.stage_bytes = 0;
  // the captured initialize() graph bakes `stage + stage_bytes` as the right image's source: a reallocation (even one
  // that returns the same address) invalidates it
  if (&lane == &h->lanes[0] && h->init_graph) {
    cudaGraphExecDestroy(h->init_graph);
    h->init_graph = nullptr;
  }
  if (&lane == &h->lanes[0]) invalidate_step_graphs(h);
  const size_t want = (size_t)h->chunk * pair_stride;
  CUDA_TRY(cudaMalloc((void**)&lane.stage, 2 * want + 64));
  lane.stage_bytes = want;
  return VSLAM_OK;
}

int upload_images(vslam_fpg* h, Lane& lane, int p0, int n, const uint8_t* left, const uint8_t* right, size_t stride,
                  size_t pair_stride) {
  const Geometry& g = h->g;
  if (linear_upload(g, n, stride, pair_stride)) {
    // Images are contiguous on the host: ONE linear copy per side at full PCIe rate (strided 2-D/3-D copies of
    // 1241-byte rows run several times slower: 47 us instead of ~15 us for one KITTI image), then a device kernel
    // re-pitches rows to the 128 B aligned layout.  A single frame takes this path too unless its rows are so widely
    // strided (a view into a much larger image) that the linear copy would move mostly padding.
    const size_t bytes = (size_t)n * pair_stride;
    int rc = ensure_stage(h, lane, bytes, pair_stride);
    if (rc) return rc;
    for (int side = 0; side < 2; ++side) {
      const uint8_t* src = (side == 0 ? left : right) + (size_t)p0 * pair_stride;
      uint8_t* dst = lane.stage + (size_t)side * lane.stage_bytes;
      CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, lane.stream));
    }
    launch_repitch(g, lane.stage, lane.stage + lane.stage_bytes, (int)stride,
                   h->b.image + (size_t)2 * p0 * g.rows * g.pitch, n, lane.stream);
    ++h->launches;
    return VSLAM_OK;
  }
  for (int side = 0; side < 2; ++side) {
    const uint8_t* src = (side == 0 ? left : right) + (size_t)p0 * pair_stride;
    uint8_t* dst = h->b.image + ((size_t)2 * p0 + side) * g.rows * g.pitch;
    if (n > 1 && pair_stride % stride == 0 && pair_stride / stride >= (size_t)g.rows) {
      // one strided 3-D copy: host [pair][row][col] -> device [pair][side][row][col]
      cudaMemcpy3DParms p;
      std::memset(&p, 0, sizeof(p));
      p.srcPtr = make_cudaPitchedPtr(const_cast<uint8_t*>(src), stride, g.cols, pair_stride / stride);
      p.dstPtr = make_cudaPitchedPtr(dst, g.pitch, g.cols, (size_t)2 * g.rows);
      p.extent = make_cudaExtent(g.cols, g.rows, n);
      p.kind = cudaMemcpyHostToDevice;
      CUDA_TRY(cudaMemcpy3DAsync(&p, lane.stream));
    } else {
      for (int i = 0; i < n; ++i)
        CUDA_TRY(cudaMemcpy2DAsync(dst + (size_t)2 * i * g.rows * g.pitch, g.pitch, src + (size_t)i * pair_stride, stride,
                                   g.cols, g.rows, cudaMemcpyHostToDevice, lane.stream));
    }
  }
  return VSLAM_OK;
}

// sorted (row, col) feature order -> the reference's order (region by region, row-major inside a region)
void reference_order(const vslam_fpg* h, const std::vector<uint32_t>& xy, std::vector<int>& sorted_to_ref,
                     std::vector<int>& ref_to_sorted) {
  const int n = (int)xy.size();
  sorted_to_ref.resize(n);
  ref_to_sorted.resize(n);
  if (h->g.n_regions == 1) {
    for (int i = 0; i < n; ++i) sorted_to_ref[i] = ref_to_sorted[i] = i;
    return;
  }
  std::vector<int> region_of(n, 0);
  for (int i = 0; i < n; ++i) {
    const int x = xy[i] & 0xffff, y = xy[i] >> 16;
    for (int r = 0; r < h->g.n_regions; ++r) {
      const HostRegion& q = h->regions[r];
      if (x >= q.x + 3 && x <= q.x + q.w - 4 && y >= q.y + 3 && y <= q.y + q.h - 4) {
        region_of[i] = r;
        break;
      }
    }
  }
  int k = 0;
  for (int r = 0; r < h->g.n_regions; ++r)
    for (int i = 0; i < n; ++i)
      if (region_of[i] == r) {
        ref_to_sorted[k] = i;
        sorted_to_ref[i] = k++;
      }
}

size_t feat_capacity_bytes(const vslam_fpg* h) { return 2 * feature_pack_bytes(h->g.cap); }
const uint8_t* feat_side(const vslam_fpg* h, int side) { return h->h_feat + (side ? feature_pack_bytes(h->h_n_desc[0]) : 0); }
const uint32_t* feat_xy(const vslam_fpg* h, int side) { return reinterpret_cast<const uint32_t*>(feat_side(h, side)); }
const uint8_t* feat_score(const vslam_fpg* h, int side) { return feat_side(h, side) + sizeof(uint32_t) * (size_t)h->h_n_desc[side]; }
const uint8_t* feat_desc(const vslam_fpg* h, int side) { return feat_side(h, side) + feature_pack_desc_offset(h->h_n_desc[side]); }

// Every host of the single-pair path asks for the features right after initialize() (frame->keypointsLeft/Right(),
// descriptorsLeft/Right(): reference frame.h:64-67).  Their download -- positions, FAST responses (computed here, not
// on the path), descriptors, both sides -- is therefore started at the end of initialize() on lane 1's stream, where it
// neither delays the kernels of track() / compute() on lane 0 nor costs the caller a round trip per array and side:
// ONE kernel packs everything (pack_features_kernel), ONE copy brings it to pinned host memory.
int prefetch_features(vslam_fpg* h) {
  cudaStream_t s = h->lanes[1].stream;
  const size_t bytes = feature_pack_bytes(h->h_n_desc[0]) + feature_pack_bytes(h->h_n_desc[1]);
  if (bytes) {
    launch_pack_features(h->g, h->b, h->d_feat, s);
    ++h->launches;
    CUDA_TRY(cudaMemcpyAsync(h->h_feat, h->d_feat, bytes, cudaMemcpyDeviceToHost, s));
  }
  CUDA_TRY(cudaEventRecord(h->feat_ev, s));
  CUDA_TRY(cudaGetLastError());
  h->feat_valid = true;
  return VSLAM_OK;
}

int fetch_xy(vslam_fpg* h, int image, int n, std::vector<uint32_t>& xy) {
  xy.resize(n);
  if (h->feat_valid && image < 2) {
    CUDA_TRY(cudaEventSynchronize(h->feat_ev));
    std::memcpy(xy.data(), feat_xy(h, image), sizeof(uint32_t) * n);
    return VSLAM_OK;
  }
  if (n) CUDA_TRY(cudaMemcpy(xy.data(), h->b.kp_xy + (size_t)image * h->g.cap, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
  return VSLAM_OK;
}

int remap_records(vslam_fpg* h, int pair, vslam_framepoint* rec, int n) {
  if (h->g.n_regions == 1 || n == 0) return VSLAM_OK;
  std::vector<uint32_t> xl, xr;
  std::vector<int> s2r_l, s2r_r, tmp;
  int rc;
  if ((rc = fetch_xy(h, 2 * pair, h->h_n_desc[2 * pair], xl))) return rc;
  if ((rc = fetch_xy(h, 2 * pair + 1, h->h_n_desc[2 * pair + 1], xr))) return rc;
  reference_order(h, xl, s2r_l, tmp);
  reference_order(h, xr, s2r_r, tmp);
  for (int i = 0; i < n; ++i) {
    if (rec[i].index_left >= 0) rec[i].index_left = s2r_l[rec[i].index_left];
    if (rec[i].index_right >= 0) rec[i].index_right = s2r_r[rec[i].index_right];
  }
  return VSLAM_OK;
}

int get_features(vslam_fpg* h, int pair, int side, vslam_keypoint* kps, uint8_t* desc, int32_t capacity, int32_t* n_out) {
  if (!h->initialized || pair < 0 || pair >= h->last_pairs) return fail(VSLAM_ERR_STATE, "no features for pair %d", pair);
  if (side != 0 && side != 1) return fail(VSLAM_ERR_INVALID_ARGUMENT, "side must be 0 or 1");
  CUDA_TRY(cudaSetDevice(h->device));
  const int image = 2 * pair + side;
  const int n = h->h_n_desc[image];
  if (n_out) *n_out = n;
  if (n > capacity) return fail(VSLAM_ERR_CAPACITY, "capacity %d < %d features", capacity, n);
  std::vector<uint32_t> xy;
  int rc;
  if ((rc = fetch_xy(h, image, n, xy))) return rc;
  std::vector<int> s2r, r2s;
  reference_order(h, xy, s2r, r2s);
  const bool staged = h->feat_valid && pair == 0;   // fetch_xy has waited for the prefetch
  if (kps && staged) {
    const uint8_t* score = feat_score(h, side);
    for (int k = 0; k < n; ++k) {
      const int i = r2s[k];
      kps[k].x = (float)(xy[i] & 0xffff);
      kps[k].y = (float)(xy[i] >> 16);
      kps[k].response = (float)score[i];
    }
  } else if (kps) {
    std::vector<uint8_t> score(n);
    if (n) {
      launch_score(h->g, h->b, image, h->lanes[0].stream);
      ++h->launches;
      CUDA_TRY(cudaMemcpyAsync(score.data(), h->b.kp_score + (size_t)image * h->g.cap, n, cudaMemcpyDeviceToHost,
                               h->lanes[0].stream));
      CUDA_TRY(cudaStreamSynchronize(h->lanes[0].stream));
    }
    for (int k = 0; k < n; ++k) {
      const int i = r2s[k];
      kps[k].x = (float)(xy[i] & 0xffff);
      kps[k].y = (float)(xy[i] >> 16);
      kps[k].response = (float)score[i];
    }
  }
  if (desc && staged) {
    const uint8_t* d = feat_desc(h, side);
    if (h->g.n_regions == 1) std::memcpy(desc, d, (size_t)n * kDescBytes);   // reference order == device order
    else
      for (int k = 0; k < n; ++k) std::memcpy(desc + (size_t)k * kDescBytes, d + (size_t)r2s[k] * kDescBytes, kDescBytes);
  } else if (desc) {
    std::vector<uint8_t> d((size_t)n * kDescBytes);
    if (n) CUDA_TRY(cudaMemcpy(d.data(), h->b.desc + (size_t)image * h->g.cap * kDescBytes, d.size(), cudaMemcpyDeviceToHost));
    for (int k = 0; k < n; ++k) std::memcpy(desc + (size_t)k * kDescBytes, d.data() + (size_t)r2s[k] * kDescBytes, kDescBytes);
  }
  return VSLAM_OK;
}

// error flag, descriptor counts and raw counts of pair 0 -> pinned host mirror
int status_download(vslam_fpg* h, Lane& lane) {
  const Geometry& g = h->g;
  if (h->max_batch == 1) {   // the three are contiguous: one copy
    CUDA_TRY(cudaMemcpyAsync(h->h_status, h->d_status, sizeof(int32_t) * (2 + 2 + 2 * g.n_regions), cudaMemcpyDeviceToHost, lane.stream));
  } else {
    CUDA_TRY(cudaMemcpyAsync(h->h_counts, h->b.raw_count, sizeof(int32_t) * 2 * g.n_regions, cudaMemcpyDeviceToHost, lane.stream));
    CUDA_TRY(cudaMemcpyAsync(h->h_status, h->d_status, sizeof(int32_t) * 4, cudaMemcpyDeviceToHost, lane.stream));   // flag + n_desc[0..1]
  }
  return VSLAM_OK;
}

double host_matching_distance(const vslam_fpg* h, int localizing, int n_left) {   // :109-125
  if (localizing) return std::min(0.1 * 256, h->cfg.maximum_matching_distance_triangulation);
  const double ratio = std::min(static_cast<double>(n_left) / h->target_keypoints, 1.0);
  return std::max(ratio * h->cfg.maximum_matching_distance_triangulation, 0.1 * 256);
}

}  // namespace

extern "C" {

int vslam_fpg_create(const vslam_fpg_config* c, int device, vslam_fpg** out) {
  if (!c || !out) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  if (c->rows < 7 || c->cols < 7 || c->cols > 65535 || c->rows > 65535)
    return fail(VSLAM_ERR_INVALID_ARGUMENT, "unsupported image size %dx%d", c->cols, c->rows);
  const int nv = c->number_of_detectors_vertical, nh = c->number_of_detectors_horizontal;
  if (nv < 1 || nh < 1 || nv * nh > kMaxRegions) return fail(VSLAM_ERR_INVALID_ARGUMENT, "unsupported detector grid %dx%d", nv, nh);
  if (c->bin_size_pixels < 1) return fail(VSLAM_ERR_INVALID_ARGUMENT, "bin_size_pixels must be positive");
  if (c->maximum_epipolar_search_offset_pixels < 0 || c->maximum_epipolar_search_offset_pixels > 100)
    return fail(VSLAM_ERR_INVALID_ARGUMENT, "maximum_epipolar_search_offset_pixels out of range");
  // stereo_framepoint_generator.cpp:26-34
  const double baseline_meters = -c->bx / c->fx;
  if (!(baseline_meters > 0))
    return fail(VSLAM_ERR_INVALID_ARGUMENT, "StereoFramePointGenerator::configure|invalid baseline (m): '%f' verify intrinsic camera parameters", baseline_meters);
  if (c->descriptor_type != VSLAM_DESCRIPTOR_ORB && c->descriptor_type != VSLAM_DESCRIPTOR_BRIEF)
    return fail(VSLAM_ERR_INVALID_ARGUMENT, "unknown descriptor_type %d", c->descriptor_type);
  const bool brief = c->descriptor_type == VSLAM_DESCRIPTOR_BRIEF;
  if (brief) {
    if (!c->brief_tests) return fail(VSLAM_ERR_INVALID_ARGUMENT, "VSLAM_DESCRIPTOR_BRIEF needs brief_tests (256 x 4 int8)");
    for (int i = 0; i < 1024; ++i)   // PATCH_SIZE / 2 of xfeatures2d/src/brief.cpp
      if (c->brief_tests[i] < -24 || c->brief_tests[i] > 24) return fail(VSLAM_ERR_INVALID_ARGUMENT, "brief_tests[%d] outside +-24", i);
  }
  int rc = require_device(device);
  if (rc) return rc;

  vslam_fpg* h = new vslam_fpg();
  h->cfg = *c;
  h->device = device;
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
  Geometry& g = h->g;
  g.rows = c->rows;
  g.cols = c->cols;
  g.pitch = (c->cols + 127) & ~127;
  g.mask_words = ((c->cols + 31) / 32 + 3) & ~3;
  g.n_regions = nv * nh;
  g.bin_size = c->bin_size_pixels;
  g.enable_binning = c->enable_keypoint_binning != 0;
  g.border = brief ? 28 : 31;   // KeyPointsFilter::runByImageBorder of the extractor: PATCH/2 + KERNEL/2 | edgeThreshold
  bin_grid(g.rows, g.cols, g.bin_size, &g.rows_bin, &g.cols_bin);
  h->target_keypoints = g.rows_bin * g.cols_bin;                    // base :308
  h->target_per_detector = h->target_keypoints / g.n_regions;       // base :312 (Count)
  g.cap = c->max_keypoints_per_image > 0 ? c->max_keypoints_per_image : std::max(4096, 4 * h->target_keypoints);
  if (g.cap > 65535) g.cap = 65535;
  h->max_batch = std::max(1, c->max_batch);
  h->n_passes = 1 + 2 * c->maximum_epipolar_search_offset_pixels;   // stereo :45-50
  detector_regions(g.rows, g.cols, nv, nh, h->regions);
  for (int i = 0; i < g.n_regions; ++i) {
    const HostRegion& q = h->regions[i];
    if (q.x < 0 || q.y < 0 || q.x + q.w > g.cols || q.y + q.h > g.rows || q.w < 7 || q.h < 7) {
      delete h;
      return fail(VSLAM_ERR_INVALID_ARGUMENT, "detector region %d out of the image", i);
    }
    h->thresholds[i] = std::rint((double)c->detector_threshold_minimum);   // base :244, :13
  }
  h->sp.fx = c->fx; h->sp.fy = c->fy; h->sp.cx = c->cx; h->sp.cy = c->cy; h->sp.bx = c->bx;
  h->sp.max_matching_distance = c->maximum_matching_distance_triangulation;
  h->sp.min_disparity = c->minimum_disparity_pixels;
  h->sp.target_keypoints = h->target_keypoints;
  h->sp.localizing = 1;
  h->out_cap = g.enable_binning ? h->target_keypoints : g.cap;

  // chunk: pairs per launch.  Large enough that the one-CTA-per-image / per-pair kernels (compact, select) fill the
  // 148 SMs; the path is instruction-bound, not HBM-bound, so L2 residency of a chunk's intermediates is secondary
  // (measured: DESIGN.md section 3, `blurred` row; profiles/prof_r1j_summary.txt)
  const size_t per_pair = (size_t)2 * g.rows * (2 * g.pitch + 4 * g.mask_words);
  int chunk = (int)std::min<size_t>(256, std::max<size_t>(1, ((size_t)512u << 20) / per_pair));
  if (const char* e = std::getenv("VSLAM_CHUNK_PAIRS")) chunk = std::max(1, atoi(e));
  if (const char* e = std::getenv("VSLAM_DESCRIBE_SUB_CHUNK")) h->describe_sub_chunk = std::max(1, atoi(e));
  h->chunk = std::min(chunk, h->max_batch);

  const size_t B = h->max_batch, I = 2 * B;
  const size_t img_bytes = (size_t)g.rows * g.pitch;
  Buffers& b = h->b;
  std::memset(&b, 0, sizeof(b));
  bool ok = true;
  auto dalloc = [&](void** p, size_t bytes) {
    if (ok && cudaMalloc(p, bytes ? bytes : 1) != cudaSuccess) ok = false;
  };
  dalloc((void**)&b.image, I * img_bytes);
  // error flag, descriptor counts and raw keypoint counts share ONE allocation (and one pinned mirror): the single-pair
  // initialize() of a max_batch == 1 handle reads all three with one copy
  dalloc((void**)&h->d_status, (2 + I + I * g.n_regions) * sizeof(int32_t));
  if (ok) {
    b.error_flag = h->d_status;
    b.n_desc = h->d_status + 2;
    b.raw_count = h->d_status + 2 + I;
  }
  dalloc((void**)&b.row_ptr, I * (g.rows + 1) * sizeof(int32_t));
  dalloc((void**)&b.kp_xy, I * g.cap * sizeof(uint32_t));
  dalloc((void**)&b.kp_score, I * g.cap);
  dalloc((void**)&b.desc, I * g.cap * kDescBytes);
  dalloc((void**)&b.match, B * g.cap * sizeof(int2));
  dalloc((void**)&b.pruned_l, B * g.cap);
  dalloc((void**)&b.consumed_r, B * g.cap);
  dalloc((void**)&b.n_out, B * 2 * sizeof(int32_t));
  dalloc((void**)&h->d_out, B * h->out_cap * sizeof(FramePointRecord));
  dalloc((void**)&h->d_matches, (size_t)g.cap * sizeof(FramePointRecord));
  dalloc((void**)&h->d_n_matches, sizeof(int32_t));
  dalloc((void**)&h->d_systems, B * 32 * sizeof(double));
  dalloc((void**)&h->d_pair_errors, B * h->out_cap * sizeof(double));
  dalloc((void**)&h->d_pair_inliers, B * h->out_cap);
  dalloc((void**)&h->track_scratch.claim_l, (size_t)g.cap * sizeof(int32_t));
  dalloc((void**)&h->track_scratch.claim_r, (size_t)g.cap * sizeof(int32_t));
  dalloc((void**)&h->track_scratch.stats, 4 * sizeof(int32_t));
  dalloc((void**)&h->d_recover_n, 4 * sizeof(int32_t));
  if (brief) {
    dalloc((void**)&h->d_brief_tests, 1024);
    if (ok && cudaMemcpy(h->d_brief_tests, c->brief_tests, 1024, cudaMemcpyHostToDevice) != cudaSuccess) ok = false;
  }
  h->cfg.brief_tests = nullptr;   // the table was copied; the caller's pointer is not kept
  if (ok && !(make_fast_tensor_map(g, b.image, (int)I, &h->image_map) && make_blur_tensor_map(g, b.image, (int)I, &h->blur_map))) {
    vslam_fpg_destroy(h);
    return fail(VSLAM_ERR_CUDA, "cuTensorMapEncodeTiled failed for the image buffer (TMA is required: no fallback)");
  }
  for (int l = 0; l < kLanes; ++l) {
    dalloc((void**)&h->lanes[l].blurred, (size_t)2 * h->chunk * img_bytes * (brief ? 2 : 1));
    dalloc((void**)&h->lanes[l].mask, (size_t)2 * h->chunk * g.rows * g.mask_words * sizeof(uint32_t));
    if (ok && cudaStreamCreateWithFlags(&h->lanes[l].stream, cudaStreamNonBlocking) != cudaSuccess) ok = false;
    if (ok && !brief && !make_blurred_tensor_map(g, h->lanes[l].blurred, 2 * h->chunk, &h->lanes[l].blurred_map)) {
      vslam_fpg_destroy(h);
      return fail(VSLAM_ERR_CUDA, "cuTensorMapEncodeTiled failed for the blurred image buffer (TMA is required: no fallback)");
    }
  }
  auto halloc = [&](void** p, size_t bytes) {
    if (ok && cudaMallocHost(p, bytes) != cudaSuccess) ok = false;
  };
  halloc((void**)&h->h_status, (2 + I + I * g.n_regions) * sizeof(int32_t));
  if (ok) {
    h->h_flag = h->h_status;
    h->h_n_desc = h->h_status + 2;
    h->h_counts = h->h_status + 2 + I;
  }
  halloc((void**)&h->h_n_out, B * 2 * sizeof(int32_t));
  halloc((void**)&h->h_out_stage, (size_t)h->out_cap * sizeof(FramePointRecord));
  halloc((void**)&h->h_thr, kMaxRegions * sizeof(int32_t));
  dalloc((void**)&h->d_thr, kMaxRegions * sizeof(int32_t));
  halloc((void**)&h->h_feat, feat_capacity_bytes(h));
  dalloc((void**)&h->d_feat, feat_capacity_bytes(h));
  if (ok && cudaEventCreateWithFlags(&h->feat_ev, cudaEventDisableTiming) != cudaSuccess) ok = false;
  halloc((void**)&h->h_systems, B * 32 * sizeof(double));
  halloc((void**)&h->h_track_stats, 4 * sizeof(int32_t));
  for (auto& e : h->clock.ev)
    if (ok && cudaEventCreate(&e) != cudaSuccess) ok = false;
  if (ok && cudaEventCreateWithFlags(&h->fork_ev, cudaEventDisableTiming) != cudaSuccess) ok = false;
  for (auto& e : h->join_ev)
    if (ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) ok = false;
  if (ok && cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking) != cudaSuccess) ok = false;
  for (auto& e : h->branch_ev)
    if (ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) ok = false;
  h->branches = std::getenv("VSLAM_NO_FRAME_BRANCHES") == nullptr;
  if (ok && cudaMemset(b.error_flag, 0, sizeof(int32_t)) != cudaSuccess) ok = false;
  if (ok && cudaMemset(b.image, 0, I * img_bytes) != cudaSuccess) ok = false;   // row padding is never uninitialised
  if (!ok) {
    const cudaError_t e = cudaGetLastError();
    vslam_fpg_destroy(h);
    return fail(VSLAM_ERR_CUDA, "device allocation failed: %s", cudaGetErrorString(e));
  }
  *h->h_flag = 0;
  *out = h;
  return VSLAM_OK;
}

int vslam_fpg_destroy(vslam_fpg* h) {
  if (!h) return VSLAM_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  Buffers& b = h->b;
  cudaFree(b.image); cudaFree(h->d_status); cudaFree(b.row_ptr); cudaFree(b.kp_xy); cudaFree(b.kp_score);
  cudaFree(b.desc); cudaFree(b.match); cudaFree(b.pruned_l); cudaFree(b.consumed_r); cudaFree(b.n_out);
  cudaFree(h->d_out); cudaFree(h->d_matches); cudaFree(h->d_n_matches); cudaFree(h->d_tracked);
  for (auto& l : h->lanes) {
    cudaFree(l.blurred);
    cudaFree(l.mask);
    cudaFree(l.stage);
    if (l.stream) cudaStreamDestroy(l.stream);
  }
  cudaFree(h->d_systems); cudaFree(h->d_pair_errors); cudaFree(h->d_pair_inliers);
  cudaFree(h->d_previous); cudaFree(h->track_scratch.tentative); cudaFree(h->track_scratch.final); cudaFree(h->track_scratch.claim_l);
  cudaFree(h->track_scratch.claim_r); cudaFree(h->track_scratch.stats); cudaFree(h->d_tracks); cudaFree(h->d_lost);
  cudaFree(h->d_recover_xy); cudaFree(h->d_recover_desc); cudaFree(h->d_recovered); cudaFree(h->d_recover_n);
  cudaFree(h->d_brief_tests);
  cudaFreeHost(h->h_track_stats); cudaFreeHost(h->h_tracks); cudaFreeHost(h->h_lost);
  if (h->init_graph) cudaGraphExecDestroy(h->init_graph);
  invalidate_step_graphs(h);
  cudaFree(h->d_step); cudaFree(h->d_step_ctl); cudaFree(h->d_step_planes); cudaFree(h->d_step_errors);
  cudaFree(h->d_step_inliers); cudaFree(h->d_step_system); cudaFree(h->d_step_track_length); cudaFree(h->d_step_kept_pos);
  cudaFree(h->d_step_bins); cudaFree(h->d_step_estimates);
  for (auto& b : h->step_inbox) cudaFree(b);
  for (auto& e : h->step_inbox_ev)
    if (e) cudaEventDestroy(e);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  cudaFreeHost(h->h_step); cudaFreeHost(h->h_step_T);
  cudaFreeHost(h->h_thr);
  cudaFree(h->d_thr);
  cudaFreeHost(h->h_feat);
  cudaFree(h->d_feat);
  cudaFreeHost(h->h_out_stage);
  if (h->feat_ev) cudaEventDestroy(h->feat_ev);
  cudaFreeHost(h->h_status); cudaFreeHost(h->h_n_out);
  cudaFreeHost(h->h_systems);
  for (auto& e : h->clock.ev)
    if (e) cudaEventDestroy(e);
  if (h->fork_ev) cudaEventDestroy(h->fork_ev);
  for (auto& e : h->join_ev)
    if (e) cudaEventDestroy(e);
  for (auto& e : h->branch_ev)
    if (e) cudaEventDestroy(e);
  if (h->side_stream) cudaStreamDestroy(h->side_stream);
  delete h;
  return VSLAM_OK;
}

int vslam_fpg_info(const vslam_fpg* h, int32_t* n_regions, int32_t* regions_xywh, int32_t* rows_bin, int32_t* cols_bin,
                   int32_t* target) {
  if (!h) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null handle");
  if (n_regions) *n_regions = h->g.n_regions;
  if (regions_xywh)
    for (int i = 0; i < h->g.n_regions; ++i) {
      regions_xywh[4 * i] = h->regions[i].x;
      regions_xywh[4 * i + 1] = h->regions[i].y;
      regions_xywh[4 * i + 2] = h->regions[i].w;
      regions_xywh[4 * i + 3] = h->regions[i].h;
    }
  if (rows_bin) *rows_bin = h->g.rows_bin;
  if (cols_bin) *cols_bin = h->g.cols_bin;
  if (target) *target = h->target_keypoints;
  return VSLAM_OK;
}

int vslam_fpg_get_thresholds(const vslam_fpg* h, double* t) {
  if (!h || !t) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  for (int i = 0; i < h->g.n_regions; ++i) t[i] = h->thresholds[i];
  return VSLAM_OK;
}

int vslam_fpg_set_thresholds(vslam_fpg* h, const double* t) {
  if (!h || !t) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  for (int i = 0; i < h->g.n_regions; ++i) h->thresholds[i] = std::rint(t[i]);
  return VSLAM_OK;
}

// host side of a finished single-pair detection: adjustDetectorThresholds (base :440-459) over the two detections
// (L, R) of this frame and the handle's frame state
static void finish_detection(vslam_fpg* h, int localizing) {
  const Geometry& g = h->g;
  for (int i = 0; i < g.n_regions; ++i) {
    double acc = 0;
    for (int side = 0; side < 2; ++side)
      acc += threshold_proposal(h->thresholds[i], h->h_counts[side * g.n_regions + i], h->target_per_detector,
                                h->cfg.target_number_of_keypoints_tolerance, h->cfg.detector_threshold_maximum_change,
                                h->cfg.detector_threshold_minimum, h->cfg.detector_threshold_maximum);
    h->thresholds[i] = std::rint(acc / 2);
  }
  h->localizing = localizing != 0;
  h->sp.localizing = h->localizing;
  h->matching_distance = host_matching_distance(h, h->localizing, h->h_n_desc[0]);
  h->initialized = true;
  h->last_pairs = 1;
  h->n_device_tracks = -1;
}

int vslam_fpg_initialize(vslam_fpg* h, const uint8_t* left, const uint8_t* right, size_t stride, int localizing,
                         int32_t* n_left, int32_t* n_right) {
  VSLAM_NVTX("vslam_fpg_initialize [KeypointDetection + DescriptorExtraction]");
  if (!h || !left || !right)   // stereo :75-78
    return fail(VSLAM_ERR_INVALID_ARGUMENT, "StereoFramePointGenerator::initialize|called with empty frame");
  if (stride < (size_t)h->g.cols) return fail(VSLAM_ERR_INVALID_ARGUMENT, "stride smaller than the image width");
  CUDA_TRY(cudaSetDevice(h->device));
  Lane& lane = h->lanes[0];
  const Geometry& g = h->g;
  // an outstanding feature prefetch (lane 1) still reads image / kp_xy / desc of the previous frame: order this
  // frame's writes (lane 0) behind it
  if (h->feat_valid) CUDA_TRY(cudaStreamWaitEvent(lane.stream, h->feat_ev, 0));
  h->feat_valid = false;
  int rc = VSLAM_OK;
  const size_t image_bytes = stride * (size_t)g.rows;
  bool launched = false;
  if (h->init_graph_ok && linear_upload(g, 1, stride, image_bytes)) {
    if ((rc = ensure_stage(h, lane, image_bytes, image_bytes))) return rc;
    if (h->init_graph && (h->init_graph_stage != lane.stage || h->init_graph_stride != stride ||
                          h->init_graph_profiling != h->profiling)) {
      cudaGraphExecDestroy(h->init_graph);
      h->init_graph = nullptr;
    }
    if (!h->init_graph) {   // capture the device side of this call once
      const int64_t launches_before = h->launches;
      cudaGraph_t graph = nullptr;
      bool ok = cudaStreamBeginCapture(lane.stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
      if (ok) {
        h->capturing = true;
        {   // the thresholds (and the mask clear of a multi-region grid) travel beside the repitch: FAST is the first
            // kernel that needs either (a copy node in front of the chain costs ~3 us of a single frame)
            const bool beside = use_branches(h, lane, 1);
            cudaStream_t s = beside ? h->side_stream : lane.stream;
            if (beside) order_after(h, lane.stream, s, 0);
            cudaMemcpyAsync(h->d_thr, h->h_thr, sizeof(int32_t) * g.n_regions, cudaMemcpyHostToDevice, s);
            if (g.n_regions > 1) cudaMemsetAsync(lane.mask, 0, (size_t)2 * g.rows * g.mask_words * sizeof(uint32_t), s);
            launch_repitch(g, lane.stage, lane.stage + lane.stage_bytes, (int)stride, h->b.image, 1, lane.stream,
                           h->b.raw_count, 2 * g.n_regions);
            ++h->launches;
            if (beside) order_after(h, s, lane.stream, 0);
        }
        run_detect_describe(h, lane, 0, 1, h->d_thr, true, true);
        status_download(h, lane);
        h->capturing = false;
        ok = cudaStreamEndCapture(lane.stream, &graph) == cudaSuccess && graph != nullptr;
      }
      if (ok) ok = cudaGraphInstantiate(&h->init_graph, graph, 0) == cudaSuccess;
      if (graph) cudaGraphDestroy(graph);
      h->init_graph_kernels = (int)(h->launches - launches_before);
      h->launches = launches_before;
      if (!ok) {            // no graph on this system / for this configuration: the plain launches below do the same work
        cudaGetLastError();
        h->init_graph = nullptr;
        h->init_graph_ok = false;
      } else {
        h->init_graph_stage = lane.stage;
        h->init_graph_stride = stride;
        h->init_graph_profiling = h->profiling;
      }
    }
    if (h->init_graph) {
      for (int i = 0; i < g.n_regions; ++i) {   // FastDetector::setThreshold -> std::rint (:21); cv::FAST clamps
        const int t = (int)std::rint(h->thresholds[i]);
        h->h_thr[i] = std::min(std::max(t, 0), 255);
      }
      CUDA_TRY(cudaMemcpyAsync(lane.stage, left, image_bytes, cudaMemcpyHostToDevice, lane.stream));
      CUDA_TRY(cudaMemcpyAsync(lane.stage + lane.stage_bytes, right, image_bytes, cudaMemcpyHostToDevice, lane.stream));
      CUDA_TRY(cudaGraphLaunch(h->init_graph, lane.stream));
      h->launches += h->init_graph_kernels;
      ++h->graph_launches;
      launched = true;
    }
  }
  if (!launched) {
    if ((rc = upload_images(h, lane, 0, 1, left, right, stride, image_bytes))) return rc;
    run_detect_describe(h, lane, 0, 1);
    if ((rc = status_download(h, lane))) return rc;
  }
  CUDA_TRY(cudaStreamSynchronize(lane.stream));
  CUDA_TRY(cudaGetLastError());
  collect_clock(h, true, false);
  if ((rc = check_flag(h))) {
    cudaMemset(h->b.error_flag, 0, sizeof(int32_t));
    return rc;
  }
  finish_detection(h, localizing);
  if (n_left) *n_left = h->h_n_desc[0];
  if (n_right) *n_right = h->h_n_desc[1];
  return prefetch_features(h);
}

int64_t vslam_fpg_graph_launch_count(const vslam_fpg* h) { return h ? h->graph_launches : 0; }

int vslam_fpg_get_features(vslam_fpg* h, int side, vslam_keypoint* kps, uint8_t* desc, int32_t capacity, int32_t* n) {
  if (!h) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null handle");
  return get_features(h, 0, side, kps, desc, capacity, n);
}

int vslam_fpg_batch_get_features(vslam_fpg* h, int32_t pair, int side, vslam_keypoint* kps, uint8_t* desc,
                                 int32_t capacity, int32_t* n) {
  if (!h) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null handle");
  return get_features(h, pair, side, kps, desc, capacity, n);
}

int vslam_fpg_get_detection_stats(vslam_fpg* h, int32_t* cl, int32_t* cr, double* matching_distance) {
  if (!h) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null handle");
  if (!h->initialized) return fail(VSLAM_ERR_STATE, "initialize has not run");
  for (int i = 0; i < h->g.n_regions; ++i) {
    if (cl) cl[i] = h->h_counts[i];
    if (cr) cr[i] = h->h_counts[h->g.n_regions + i];
  }
  if (matching_distance) *matching_distance = h->matching_distance;
  return VSLAM_OK;
}

int vslam_fpg_compute(vslam_fpg* h, const vslam_tracked_point* tracked, int32_t n_tracked, vslam_framepoint* out,
                      int32_t capacity, int32_t* n_out, int32_t* n_matches) {
  VSLAM_NVTX("vslam_fpg_compute [StereoMatching]");
  if (!h) return fail(VSLAM_ERR_INVALID_ARGUMENT, "StereoFramePointGenerator::compute|called with empty frame");
  if (!h->initialized || h->last_pairs != 1) return fail(VSLAM_ERR_STATE, "compute without initialize");
  const bool device_tracks = n_tracked == VSLAM_TRACKED_FROM_LAST_TRACK;
  if (device_tracks) {
    if (h->n_device_tracks < 0) return fail(VSLAM_ERR_STATE, "VSLAM_TRACKED_FROM_LAST_TRACK without vslam_fpg_track on this frame");
    n_tracked = h->n_device_tracks;
    tracked = nullptr;
  } else if (n_tracked < 0 || (n_tracked > 0 && !tracked)) {
    return fail(VSLAM_ERR_INVALID_ARGUMENT, "bad tracked points");
  }
  CUDA_TRY(cudaSetDevice(h->device));
  Lane& lane = h->lanes[0];
  if (!device_tracks) h->n_device_tracks = -1;   // d_tracked is about to be overwritten
  if (!device_tracks && n_tracked > h->tracked_cap) {
    invalidate_step_graphs(h);
    cudaFree(h->d_tracked);
    h->d_tracked = nullptr;
    h->tracked_cap = 0;
    CUDA_TRY(cudaMalloc((void**)&h->d_tracked, sizeof(TrackedPoint) * (size_t)n_tracked * 2));
    h->tracked_cap = n_tracked * 2;
  }
  if (n_tracked && !device_tracks)
    CUDA_TRY(cudaMemcpyAsync(h->d_tracked, tracked, sizeof(TrackedPoint) * n_tracked, cudaMemcpyHostToDevice, lane.stream));
  // the strip kernel keeps bin state in float: exact for everything the stereo path produces (disparities are float
  // differences, distances Hamming counts); other values take the generic double-precision kernel
  bool generic = false;
  for (int32_t i = 0; i < n_tracked && !generic && !device_tracks; ++i)
    generic = (double)(float)tracked[i].disparity != tracked[i].disparity ||
              (double)(float)tracked[i].distance != tracked[i].distance || tracked[i].distance < 0 ||
              tracked[i].row < 0 || tracked[i].col < 0;
  run_match_select(h, lane, 0, 1, h->d_tracked, n_tracked, generic);
  const FramePointRecord* src = h->d_out;
  if (!h->g.enable_binning) {   // :456-460 : every new point, in emission order
    launch_emit_matches(h->g, h->sp, h->b, 0, h->n_passes, h->d_matches, h->g.cap, h->d_n_matches, lane.stream);
    ++h->launches;
    CUDA_TRY(cudaMemcpyAsync(h->h_n_out, h->d_n_matches, sizeof(int32_t), cudaMemcpyDeviceToHost, lane.stream));
    CUDA_TRY(cudaMemcpyAsync(h->h_n_out + 1, h->b.n_out + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, lane.stream));
    src = h->d_matches;
  } else {
    CUDA_TRY(cudaMemcpyAsync(h->h_n_out, h->b.n_out, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, lane.stream));
  }
  CUDA_TRY(cudaMemcpyAsync(h->h_flag, h->b.error_flag, sizeof(int32_t), cudaMemcpyDeviceToHost, lane.stream));
  // the records travel with the counts (one round trip instead of two): the whole bin-sized buffer when it is small
  const bool staged_records = h->g.enable_binning && (size_t)h->out_cap * sizeof(FramePointRecord) <= (256u << 10);
  if (staged_records)
    CUDA_TRY(cudaMemcpyAsync(h->h_out_stage, src, sizeof(FramePointRecord) * (size_t)h->out_cap, cudaMemcpyDeviceToHost, lane.stream));
  CUDA_TRY(cudaStreamSynchronize(lane.stream));
  CUDA_TRY(cudaGetLastError());
  collect_clock(h, false, true);
  int rc = check_flag(h);
  if (rc) {
    cudaMemset(h->b.error_flag, 0, sizeof(int32_t));
    return rc;
  }
  const int n = h->h_n_out[0];
  if (n_out) *n_out = n;
  if (n_matches) *n_matches = h->h_n_out[1];
  if (n > capacity) return fail(VSLAM_ERR_CAPACITY, "capacity %d < %d framepoints", capacity, n);
  if (n && !out) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null output");
  if (n && staged_records) std::memcpy(out, h->h_out_stage, sizeof(FramePointRecord) * n);
  else if (n) CUDA_TRY(cudaMemcpy(out, src, sizeof(FramePointRecord) * n, cudaMemcpyDeviceToHost));
  return remap_records(h, 0, out, n);
}

// grows the per-previous-point buffers of track() / recoverPoints()
static int ensure_previous_capacity(vslam_fpg* h, int n) {
  if (n <= h->previous_cap) return VSLAM_OK;
  CUDA_TRY(cudaStreamSynchronize(h->lanes[0].stream));
  invalidate_step_graphs(h);
  const int cap = std::max(2 * n, 1024);
  cudaFree(h->d_previous); cudaFree(h->track_scratch.tentative); cudaFree(h->track_scratch.final); cudaFree(h->d_tracks); cudaFree(h->d_lost);
  cudaFree(h->d_recover_xy); cudaFree(h->d_recover_desc); cudaFree(h->d_recovered);
  h->d_previous = nullptr; h->track_scratch.tentative = nullptr; h->track_scratch.final = nullptr; h->d_tracks = nullptr; h->d_lost = nullptr;
  h->d_recover_xy = nullptr; h->d_recover_desc = nullptr; h->d_recovered = nullptr;
  cudaFreeHost(h->h_tracks); cudaFreeHost(h->h_lost);
  h->h_tracks = nullptr; h->h_lost = nullptr;
  h->previous_cap = 0;
  CUDA_TRY(cudaMalloc((void**)&h->d_previous, sizeof(PreviousPoint) * (size_t)cap));
  CUDA_TRY(cudaMalloc((void**)&h->track_scratch.tentative, sizeof(int4) * (size_t)cap));
  CUDA_TRY(cudaMalloc((void**)&h->track_scratch.final, sizeof(int4) * (size_t)cap));
  CUDA_TRY(cudaMalloc((void**)&h->d_tracks, sizeof(TrackRecord) * (size_t)cap));
  CUDA_TRY(cudaMalloc((void**)&h->d_lost, sizeof(int32_t) * (size_t)cap));
  CUDA_TRY(cudaMalloc((void**)&h->d_recover_xy, sizeof(uint32_t) * 2 * (size_t)cap));
  CUDA_TRY(cudaMalloc((void**)&h->d_recover_desc, (size_t)cap * (2 * kDescBytes + 1)));
  CUDA_TRY(cudaMalloc((void**)&h->d_recovered, sizeof(RecoveredRecord) * (size_t)cap));
  CUDA_TRY(cudaMallocHost((void**)&h->h_tracks, sizeof(TrackRecord) * (size_t)cap));
  CUDA_TRY(cudaMallocHost((void**)&h->h_lost, sizeof(int32_t) * (size_t)cap));
  h->previous_cap = cap;
  if (cap > h->tracked_cap) {
    cudaFree(h->d_tracked);
    h->d_tracked = nullptr;
    h->tracked_cap = 0;
    CUDA_TRY(cudaMalloc((void**)&h->d_tracked, sizeof(TrackedPoint) * (size_t)cap));
    h->tracked_cap = cap;
  }
  return VSLAM_OK;
}

int vslam_fpg_track(vslam_fpg* h, const vslam_previous_point* previous, int32_t n_previous, const double T[12],
                    int track_by_appearance, int32_t projection_tracking_distance_pixels,
                    double maximum_descriptor_distance_tracking, vslam_track* tracks, int32_t capacity,
                    int32_t* n_tracks, int32_t* lost, int32_t* n_lost, int32_t* n_tracked_landmarks,
                    double* average_descriptor_distance) {
  VSLAM_NVTX("vslam_fpg_track [PoseTracker3D::compute->track]");
  if (!h) return fail(VSLAM_ERR_INVALID_ARGUMENT, "StereoFramePointGenerator::track|called with invalid frames");   // :468-471
  if (!T || n_previous < 0 || (n_previous > 0 && !previous)) return fail(VSLAM_ERR_INVALID_ARGUMENT, "bad previous points / transform");
  if (!h->initialized || h->last_pairs != 1) return fail(VSLAM_ERR_STATE, "track without initialize");
  if (projection_tracking_distance_pixels < 0) return fail(VSLAM_ERR_INVALID_ARGUMENT, "negative tracking distance");
  CUDA_TRY(cudaSetDevice(h->device));
  Lane& lane = h->lanes[0];
  int rc = ensure_previous_capacity(h, n_previous);
  if (rc) return rc;
  if (n_previous)
    CUDA_TRY(cudaMemcpyAsync(h->d_previous, previous, sizeof(PreviousPoint) * (size_t)n_previous, cudaMemcpyHostToDevice, lane.stream));
  TrackParams tp;
  for (int i = 0; i < 12; ++i) tp.T[i] = T[i];
  tp.by_appearance = track_by_appearance != 0;
  tp.distance_pixels = projection_tracking_distance_pixels;
  tp.max_distance_tracking = maximum_descriptor_distance_tracking;
  mark(h, lane, kEvTrack0);
  launch_track(h->g, h->sp, h->b, 0, h->d_previous, n_previous, tp, h->track_scratch, h->d_tracks, h->d_lost,
               h->d_tracked, lane.stream);
  mark(h, lane, kEvTrack1);
  h->launches += n_previous > 0 ? 3 : 1;
  // results travel with the counts in ONE round trip: at most n_previous records each (ordered, valid prefix)
  CUDA_TRY(cudaMemcpyAsync(h->h_track_stats, h->track_scratch.stats, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, lane.stream));
  if (n_previous) {
    CUDA_TRY(cudaMemcpyAsync(h->h_tracks, h->d_tracks, sizeof(TrackRecord) * (size_t)n_previous, cudaMemcpyDeviceToHost, lane.stream));
    CUDA_TRY(cudaMemcpyAsync(h->h_lost, h->d_lost, sizeof(int32_t) * (size_t)n_previous, cudaMemcpyDeviceToHost, lane.stream));
  }
  CUDA_TRY(cudaStreamSynchronize(lane.stream));
  CUDA_TRY(cudaGetLastError());
  if (h->profiling) add_interval(h, kKTrack, kEvTrack0, kEvTrack1, n_previous > 0 ? 3 : 1);
  const int nt = h->h_track_stats[0], nl = h->h_track_stats[1];
  h->n_device_tracks = nt;
  if (n_tracks) *n_tracks = nt;
  if (n_lost) *n_lost = nl;
  if (n_tracked_landmarks) *n_tracked_landmarks = h->h_track_stats[2];
  if (average_descriptor_distance)   // :666-667 (0/0 -> NaN like the reference)
    *average_descriptor_distance = nt ? (double)h->h_track_stats[3] / (double)nt : std::nan("");
  if (nt > capacity) return fail(VSLAM_ERR_CAPACITY, "capacity %d < %d tracks", capacity, nt);
  if (nt && !tracks) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null output");
  if (nt) std::memcpy(tracks, h->h_tracks, sizeof(TrackRecord) * (size_t)nt);
  if (nl && lost) std::memcpy(lost, h->h_lost, sizeof(int32_t) * (size_t)nl);
  if (h->g.n_regions > 1 && nt) {   // sorted device order -> the reference's keypoint order
    std::vector<uint32_t> xl, xr;
    std::vector<int> s2r_l, s2r_r, tmp;
    if ((rc = fetch_xy(h, 0, h->h_n_desc[0], xl))) return rc;
    if ((rc = fetch_xy(h, 1, h->h_n_desc[1], xr))) return rc;
    reference_order(h, xl, s2r_l, tmp);
    reference_order(h, xr, s2r_r, tmp);
    for (int i = 0; i < nt; ++i) {
      tracks[i].index_left = s2r_l[tracks[i].index_left];
      tracks[i].index_right = s2r_r[tracks[i].index_right];
    }
  }
  return VSLAM_OK;
}

int vslam_fpg_prune_tracks(vslam_fpg* h, vslam_aligner* aligner, double maximum_error_kernel, int32_t* n_kept,
                           uint8_t* kept) {
  VSLAM_NVTX("vslam_fpg_prune_tracks [PoseTracker3D::_prunePoints]");
  if (!h || !aligner) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null handle");
  if (h->n_device_tracks < 0) return fail(VSLAM_ERR_STATE, "no tracks of vslam_fpg_track on the device");
  AlignerResults r;
  int rc = fetch_aligner_results(aligner, &r);
  if (rc) return rc;
  if (r.device != h->device) return fail(VSLAM_ERR_INVALID_ARGUMENT, "generator and aligner live on different devices");
  const int n = h->n_device_tracks;
  if (r.n != n) return fail(VSLAM_ERR_STATE, "the aligner holds %d correspondences, the last track() produced %d", r.n, n);
  if (n > 8192) return fail(VSLAM_ERR_CAPACITY, "at most 8192 tracks can be pruned on the device");
  // the branch of pose_tracker_3d.cpp:441 on the host (averageError = total error / correspondences, base_aligner.h:46),
  // the per-point test and the ordered compaction on the device; the count comes from the host copy of the same flags
  const bool inliers_only = n > 0 && r.total_error / n < maximum_error_kernel;
  const double cap = 100 * maximum_error_kernel;
  int count = 0;
  for (int k = 0; k < n; ++k) {
    const bool keep = inliers_only ? r.h_inliers[k] != 0 : (r.h_errors[k] != -1.0 && r.h_errors[k] < cap);
    if (kept) kept[k] = keep;
    count += keep;
  }
  CUDA_TRY(cudaSetDevice(h->device));
  launch_prune_tracked(h->d_tracked, n, r.d_errors, r.d_inliers, inliers_only, cap, h->lanes[0].stream);
  ++h->launches;
  CUDA_TRY(cudaGetLastError());
  h->n_device_tracks = count;
  if (n_kept) *n_kept = count;
  return VSLAM_OK;
}

// ---- fused tracked frame ------------------------------------------------------------------------------------------

static int ensure_previous_capacity(vslam_fpg* h, int n);

static int setup_frame_step(vslam_fpg* h) {
  if (h->step_cap > 0) return VSLAM_OK;
  if (h->step_cap < 0) return fail(VSLAM_ERR_STATE, "the fused frame is not available on this device (no thread-block cluster of 8 CTAs)");
  if (!h->g.enable_binning) return fail(VSLAM_ERR_STATE, "vslam_fpg_frame_step needs enable_keypoint_binning");
  CUDA_TRY(cudaSetDevice(h->device));
  const int blocks = frame_step_cluster_blocks();
  if (blocks < 8) {
    h->step_cap = -1;
    return fail(VSLAM_ERR_STATE, "the fused frame is not available on this device (no thread-block cluster of 8 CTAs)");
  }
  const int cap = blocks * 256;
  int rc = ensure_previous_capacity(h, cap);
  if (rc) return rc;
  const size_t C = cap;
  CUDA_TRY(cudaMalloc((void**)&h->d_step, sizeof(FrameStepState)));
  CUDA_TRY(cudaMemset(h->d_step, 0, sizeof(FrameStepState)));
  CUDA_TRY(cudaMalloc((void**)&h->d_step_ctl, sizeof(GnControl)));
  CUDA_TRY(cudaMalloc((void**)&h->d_step_planes, sizeof(double) * 9 * C));
  CUDA_TRY(cudaMalloc((void**)&h->d_step_errors, sizeof(double) * C));
  CUDA_TRY(cudaMalloc((void**)&h->d_step_inliers, C));
  CUDA_TRY(cudaMalloc((void**)&h->d_step_system, sizeof(double) * 32));
  CUDA_TRY(cudaMalloc((void**)&h->d_step_track_length, sizeof(int32_t) * C));
  CUDA_TRY(cudaMalloc((void**)&h->d_step_kept_pos, sizeof(int32_t) * C));
  CUDA_TRY(cudaMalloc((void**)&h->d_step_bins, sizeof(int32_t) * ((size_t)h->g.rows_bin * h->g.cols_bin + 1)));
  CUDA_TRY(cudaMalloc((void**)&h->d_step_estimates, sizeof(LandmarkEstimate) * C));
  CUDA_TRY(cudaMemset(h->d_step_estimates, 0, sizeof(LandmarkEstimate) * C));
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  size_t off = up(sizeof(FrameStepHeader));
  h->step_off_tracks = off;        off = up(off + sizeof(TrackRecord) * C);
  h->step_off_kept = off;          off = up(off + C);
  h->step_off_errors = off;        off = up(off + sizeof(double) * C);
  h->step_off_inliers = off;       off = up(off + C);
  h->step_off_lost = off;          off = up(off + sizeof(int32_t) * C);
  h->step_off_points = off;        off = up(off + sizeof(FramePointRecord) * (size_t)h->out_cap);
  h->step_off_frame_points = off;  off = up(off + sizeof(PreviousPoint) * C);
  CUDA_TRY(cudaHostAlloc((void**)&h->h_step, off, cudaHostAllocMapped));
  std::memset(h->h_step, 0, off);
  CUDA_TRY(cudaHostGetDevicePointer((void**)&h->h_step_device, h->h_step, 0));
  CUDA_TRY(cudaMallocHost((void**)&h->h_step_T, offsetof(FrameStepState, n_previous)));   // the copied prefix of FrameStepState
  std::memset(h->h_step_T, 0, offsetof(FrameStepState, n_previous));
  h->step_poll = std::getenv("VSLAM_FRAME_STEP_SYNC") == nullptr;
  CUDA_TRY(cudaHostGetDevicePointer((void**)&h->h_status_device, h->h_status, 0));
  h->step_cluster_blocks = blocks;
  h->step_cap = cap;
  return VSLAM_OK;
}

static FrameStepBuffers frame_step_buffers(vslam_fpg* h) {
  FrameStepBuffers f;
  const size_t C = h->step_cap;
  f.state = h->d_step;
  f.previous = h->d_previous;
  f.tracks = h->d_tracks;
  f.lost = h->d_lost;
  f.tracked = h->d_tracked;
  f.stats = h->track_scratch.stats;
  f.track_length = h->d_step_track_length;
  f.kept_pos = h->d_step_kept_pos;
  f.points = h->d_out;
  f.n_out = h->b.n_out;
  f.error_flag = h->b.error_flag;
  f.desc = h->b.desc;
  f.ctl = h->d_step_ctl;
  f.aligner.moving = h->d_step_planes;
  f.aligner.fixed = h->d_step_planes + 3 * C;
  f.aligner.omega = h->d_step_planes + 7 * C;
  f.aligner.wt = h->d_step_planes + 8 * C;
  f.aligner.errors = h->d_step_errors;
  f.aligner.inliers = h->d_step_inliers;
  f.aligner.partials = nullptr;    // (the cluster kernel keeps its partials in shared memory)
  f.aligner.system = h->d_step_system;
  f.aligner.ticket = nullptr;
  f.aligner.stride = h->step_cap;
  f.cap = h->step_cap;
  f.out_cap = h->out_cap;
  uint8_t* d = h->h_step_device;
  f.h_header = reinterpret_cast<FrameStepHeader*>(d);
  f.h_tracks = reinterpret_cast<TrackRecord*>(d + h->step_off_tracks);
  f.h_kept = d + h->step_off_kept;
  f.h_errors = reinterpret_cast<double*>(d + h->step_off_errors);
  f.h_inliers = d + h->step_off_inliers;
  f.h_lost = reinterpret_cast<int32_t*>(d + h->step_off_lost);
  f.h_points = reinterpret_cast<FramePointRecord*>(d + h->step_off_points);
  f.h_frame_points = reinterpret_cast<PreviousPoint*>(d + h->step_off_frame_points);
  f.estimates = h->d_step_estimates;
  f.d_status = h->d_status;
  f.d_raw_count = h->b.raw_count;
  f.h_status = h->h_status_device;
  f.h_counts = h->h_status_device + (h->h_counts - h->h_status);
  f.n_regions = h->g.n_regions;
  return f;
}

// the device side of one tracked frame on lane 0's stream (captured once per frame status, or issued directly)
static int issue_frame_step(vslam_fpg* h, Lane& lane, size_t stride, const vslam_frame_step_parameters& p,
                            const uint8_t* image_left, const uint8_t* image_right) {
  const Geometry& g = h->g;
  {   // the thresholds and the motion prior travel beside the repitch (FAST is the first kernel that reads either)
    const bool beside = use_branches(h, lane, 1);
    cudaStream_t s = beside ? h->side_stream : lane.stream;
    if (beside) order_after(h, lane.stream, s, 0);
    // ONE copy node: T_prior, frame_id, ticket = 0 and the thresholds are the leading members of FrameStepState
    CUDA_TRY(cudaMemcpyAsync(h->d_step, h->h_step_T, offsetof(FrameStepState, thresholds) + sizeof(int32_t) * g.n_regions,
                             cudaMemcpyHostToDevice, s));
    // several detector regions OR their keypoints into the mask: it is cleared here, beside the repitch, not in front of FAST
    if (g.n_regions > 1)
      CUDA_TRY(cudaMemsetAsync(lane.mask, 0, (size_t)2 * g.rows * g.mask_words * sizeof(uint32_t), s));
    launch_repitch(g, image_left, image_right, (int)stride, h->b.image, 1, lane.stream, h->b.raw_count, 2 * g.n_regions);
    ++h->launches;
    if (beside) order_after(h, s, lane.stream, 0);
  }
  run_detect_describe(h, lane, 0, 1, h->d_step->thresholds, true, true);         // pose_tracker_3d.cpp:80
  const FrameStepBuffers f = frame_step_buffers(h);
  FrameStepParams fp;
  fp.max_reliable_depth = p.maximum_reliable_depth_meters;
  fp.inverse_depth_weight = p.enable_inverse_depth_as_information != 0;
  fp.error_kernel = p.aligner.maximum_error_kernel;
  fp.min_track_length = p.minimum_track_length_for_landmark_creation;
  fp.publish_frame_points = p.publish_frame_points != 0;
  TrackParams tp;
  for (int i = 0; i < 12; ++i) tp.T[i] = 0;                                      // (read from FrameStepState)
  tp.by_appearance = p.track_by_appearance != 0;
  tp.distance_pixels = p.projection_tracking_distance_pixels;
  tp.max_distance_tracking = p.maximum_descriptor_distance_tracking;
  mark(h, lane, kEvTrack0);
  launch_track(g, h->sp, h->b, 0, h->d_previous, h->step_cap, tp, h->track_scratch, h->d_tracks, h->d_lost, h->d_tracked,
               lane.stream, h->d_step);                                          // :239
  mark(h, lane, kEvTrack1);
  // After track() the frame splits into two chains that meet at the bin selection: the aligner (converge, with
  // _prunePoints as its tail) and the epipolar match of the features track() left over (:210; it reads neither the tracks nor the pose).
  // The tracks' share of points() needs the prune only and runs beside the selection.
  const bool branches = use_branches(h, lane, 1);
  cudaStream_t side = branches ? h->side_stream : lane.stream;
  if (branches) order_after(h, lane.stream, side, 1);
  auto match_passes = [&](cudaStream_t s) {
    for (int pass = 0; pass < h->n_passes; ++pass) {                             // :210
      const int offset = pass == 0 ? 0 : ((pass & 1) ? (pass + 1) / 2 : -(pass / 2));
      launch_match(g, h->sp, h->b, 0, 1, pass, offset, s);
      ++h->launches;
    }
  };
  // ... and so does the replay of the matches over the bins: every pre-loaded point of a fused frame has previous(), so
  // a surviving track simply empties its bin afterwards (select_strips_kernel, kSelectMerge)
  const bool split_select = branches && select_strips_available(g);
  if (branches) match_passes(side);
  if (split_select) {
    launch_select(g, h->sp, h->b, 0, 1, h->n_passes, nullptr, 0, h->d_out, h->out_cap, false, side, nullptr, kSelectReplay,
                  h->d_step_bins);
    ++h->launches;
  }
  AlignerCamera cam;
  const double K[9] = {h->sp.fx, 0, h->sp.cx, 0, h->sp.fy, h->sp.cy, 0, 0, 1};
  for (int i = 0; i < 9; ++i) cam.K[i] = K[i];
  cam.baseline[0] = h->sp.bx; cam.baseline[1] = 0; cam.baseline[2] = 0;
  cam.rows = g.rows; cam.cols = g.cols;
  cam.min_depth = p.minimum_reliable_depth_meters;
  GnParams gp;
  gp.error_delta = p.aligner.error_delta_for_convergence;
  gp.kernel = p.aligner.maximum_error_kernel;
  gp.damping = p.aligner.damping;
  gp.max_iterations = p.aligner.maximum_number_of_iterations;
  gp.inlier_gate = p.aligner.minimum_number_of_inliers;                          // stereouv_aligner.cpp:224
  // StereoUVAligner::initialize (:124-126 / :355-356) is the head of the kernel, _prunePoints (:437-472) its tail
  const FrameFill fill = {h->d_tracks, h->d_previous, h->d_step_estimates, h->d_step_track_length, h->d_step,
                          fp.max_reliable_depth, fp.inverse_depth_weight};
  const FramePrune prune = {h->d_tracked, h->d_step_kept_pos, h->d_step, fp.error_kernel};
  CUDA_TRY(launch_converge_frame(f.aligner, cam, gp, h->d_step_ctl, h->track_scratch.stats, h->step_cluster_blocks, fill,
                                 prune, lane.stream));                           // :357
  if (branches) {
    order_after(h, side, lane.stream, 2);      // the selection needs the matches ...
    order_after(h, lane.stream, side, 3);      // ... and the tracks' share of points() the prune
    launch_frame_assemble(g, f, fp, kAssembleTracks, side);
  }
  mark(h, lane, kEvMatch0);
  if (!branches) match_passes(lane.stream);
  mark(h, lane, kEvMatch1);
  launch_select(g, h->sp, h->b, 0, 1, h->n_passes, h->d_tracked, 0, h->d_out, h->out_cap, false, lane.stream,
                &h->d_step->n_kept, split_select ? kSelectMerge : kSelectAll, h->d_step_bins);
  mark(h, lane, kEvSelect1);
  if (branches) {
    order_after(h, side, lane.stream, 1);
    launch_frame_assemble(g, f, fp, kAssembleRest, lane.stream);
    ++h->launches;
  } else {
    launch_frame_assemble(g, f, fp, kAssembleAll, lane.stream);
  }
  h->launches += 3 + 2 + 1;   // track (search, resolve, emit), converge (+ initialize, prune) / assemble, select
  return VSLAM_OK;                // (the detection status reaches the host through frame_assemble_kernel's last block)
}

int32_t vslam_fpg_frame_step_capacity(vslam_fpg* h) {
  if (!h) return 0;
  return setup_frame_step(h) == VSLAM_OK ? h->step_cap : 0;
}

int vslam_fpg_frame_step_reset(vslam_fpg* h) {
  if (!h) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null handle");
  int rc = setup_frame_step(h);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemsetAsync(h->d_step, 0, sizeof(FrameStepState), h->lanes[0].stream));
  if (h->copy_stream) CUDA_TRY(cudaStreamSynchronize(h->copy_stream));   // a new sequence: staged frames are dropped
  h->step_inbox_count = 0;
  h->step_points = 0;
  return VSLAM_OK;
}

int vslam_fpg_frame_step_set_previous(vslam_fpg* h, const vslam_previous_point* previous, int32_t n_previous) {
  if (!h || n_previous < 0 || (n_previous > 0 && !previous)) return fail(VSLAM_ERR_INVALID_ARGUMENT, "bad previous points");
  int rc = setup_frame_step(h);
  if (rc) return rc;
  if (n_previous > h->step_cap) return fail(VSLAM_ERR_CAPACITY, "%d previous points, the fused frame holds %d", n_previous, h->step_cap);
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = h->lanes[0].stream;
  if (n_previous)
    CUDA_TRY(cudaMemcpyAsync(h->d_previous, previous, sizeof(PreviousPoint) * (size_t)n_previous, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(&h->d_step->n_previous, &n_previous, sizeof(int32_t), cudaMemcpyHostToDevice, s));
  if (n_previous) CUDA_TRY(cudaMemsetAsync(h->d_step_estimates, 0, sizeof(LandmarkEstimate) * (size_t)n_previous, s));
  CUDA_TRY(cudaStreamSynchronize(s));   // the caller's buffer and the stack variable may go away
  h->step_points = n_previous;
  return VSLAM_OK;
}

int vslam_fpg_frame_step_set_landmark_estimates(vslam_fpg* h, const vslam_landmark_estimate* estimates, int32_t n_points) {
  if (!h || n_points < 0 || (n_points > 0 && !estimates)) return fail(VSLAM_ERR_INVALID_ARGUMENT, "bad landmark estimates");
  int rc = setup_frame_step(h);
  if (rc) return rc;
  if (n_points != h->step_points)
    return fail(VSLAM_ERR_STATE, "%d landmark estimates for the %d points the device holds (one entry per point of the last frame's points())",
                n_points, h->step_points);
  if (n_points == 0) return VSLAM_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = h->lanes[0].stream;
  static_assert(sizeof(vslam_landmark_estimate) == sizeof(LandmarkEstimate), "vslam_landmark_estimate layout");
  CUDA_TRY(cudaMemcpyAsync(h->d_step_estimates, estimates, sizeof(LandmarkEstimate) * (size_t)n_points, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaStreamSynchronize(s));   // the caller's buffer may go away
  return VSLAM_OK;
}

// the two inbox buffers of the fused frame, sized for `image_bytes` per image
static int ensure_inbox(vslam_fpg* h, size_t image_bytes) {
  if (h->step_inbox_bytes >= image_bytes) return VSLAM_OK;
  if (h->step_inbox_count) return fail(VSLAM_ERR_STATE, "a larger frame than the prefetched one: consume the staged pair first");
  CUDA_TRY(cudaStreamSynchronize(h->lanes[0].stream));
  if (h->copy_stream) CUDA_TRY(cudaStreamSynchronize(h->copy_stream));
  invalidate_step_graphs(h);
  for (auto& b : h->step_inbox) {
    cudaFree(b);
    b = nullptr;
  }
  h->step_inbox_bytes = 0;
  if (!h->copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  for (auto& e : h->step_inbox_ev)
    if (!e) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& b : h->step_inbox) CUDA_TRY(cudaMalloc((void**)&b, 2 * image_bytes + 64));   // (+ the repitch kernel's read slack)
  h->step_inbox_bytes = image_bytes;
  return VSLAM_OK;
}

// the two images of a frame into inbox buffer `buf`: ONE copy when the host keeps the pair back to back (left, then
// right) and the inbox is sized for exactly this image, else one per side
static int upload_pair(vslam_fpg* h, int buf, const uint8_t* left, const uint8_t* right, size_t image_bytes,
                       cudaStream_t stream) {
  if (right == left + image_bytes && h->step_inbox_bytes == image_bytes) {
    CUDA_TRY(cudaMemcpyAsync(h->step_inbox[buf], left, 2 * image_bytes, cudaMemcpyHostToDevice, stream));
    return VSLAM_OK;
  }
  CUDA_TRY(cudaMemcpyAsync(h->step_inbox[buf], left, image_bytes, cudaMemcpyHostToDevice, stream));
  CUDA_TRY(cudaMemcpyAsync(h->step_inbox[buf] + h->step_inbox_bytes, right, image_bytes, cudaMemcpyHostToDevice, stream));
  return VSLAM_OK;
}

int vslam_fpg_frame_step_prefetch(vslam_fpg* h, const uint8_t* left, const uint8_t* right, size_t stride) {
  VSLAM_NVTX("vslam_fpg_frame_step_prefetch");
  if (!h || !left || !right) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  if (stride < (size_t)h->g.cols) return fail(VSLAM_ERR_INVALID_ARGUMENT, "stride smaller than the image width");
  int rc = setup_frame_step(h);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(h->device));
  const size_t image_bytes = stride * (size_t)h->g.rows;
  if (!linear_upload(h->g, 1, stride, image_bytes)) return fail(VSLAM_ERR_INVALID_ARGUMENT, "the fused frame takes dense images (row stride close to the width)");
  if (h->step_inbox_count >= 2) return fail(VSLAM_ERR_STATE, "two frames are staged already: call vslam_fpg_frame_step");
  if ((rc = ensure_inbox(h, image_bytes))) return rc;
  // the buffer behind the staged ones: its last reader is a frame that has returned (the call is synchronous)
  const int buf = (h->step_inbox_head + h->step_inbox_count) & 1;
  if ((rc = upload_pair(h, buf, left, right, image_bytes, h->copy_stream))) return rc;
  CUDA_TRY(cudaEventRecord(h->step_inbox_ev[buf], h->copy_stream));
  h->step_inbox_stride[buf] = stride;
  ++h->step_inbox_count;
  return VSLAM_OK;
}

int vslam_fpg_frame_step(vslam_fpg* h, const uint8_t* left, const uint8_t* right, size_t stride, int localizing,
                         const double T_prior[12], const vslam_frame_step_parameters* p, vslam_frame_step_result* out) {
  VSLAM_NVTX("vslam_fpg_frame_step [PoseTracker3D::compute]");
  if (!h || !T_prior || !p || !out || (left == nullptr) != (right == nullptr)) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  const bool staged = left == nullptr;   // the oldest pair of vslam_fpg_frame_step_prefetch
  if (staged && h->step_inbox_count == 0) return fail(VSLAM_ERR_STATE, "no images and no prefetched frame");
  if (!staged && h->step_inbox_count) return fail(VSLAM_ERR_STATE, "a prefetched frame is waiting: pass null images to consume it");
  if (staged) stride = h->step_inbox_stride[h->step_inbox_head];
  if (stride < (size_t)h->g.cols) return fail(VSLAM_ERR_INVALID_ARGUMENT, "stride smaller than the image width");
  if (p->projection_tracking_distance_pixels < 0) return fail(VSLAM_ERR_INVALID_ARGUMENT, "negative tracking distance");
  if (p->aligner.maximum_number_of_iterations < 1) return fail(VSLAM_ERR_INVALID_ARGUMENT, "maximum_number_of_iterations < 1");
  int rc = setup_frame_step(h);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(h->device));
  Lane& lane = h->lanes[0];
  const Geometry& g = h->g;
  if (h->feat_valid) CUDA_TRY(cudaStreamWaitEvent(lane.stream, h->feat_ev, 0));
  h->feat_valid = false;
  const size_t image_bytes = stride * (size_t)g.rows;
  if (!linear_upload(g, 1, stride, image_bytes)) return fail(VSLAM_ERR_INVALID_ARGUMENT, "the fused frame takes dense images (row stride close to the width)");
  if ((rc = ensure_inbox(h, image_bytes))) return rc;
  const int buf = h->step_inbox_head;
  const uint8_t* image_left = h->step_inbox[buf];
  const uint8_t* image_right = h->step_inbox[buf] + h->step_inbox_bytes;
  const int L = localizing != 0;
  if (h->step_graph[0][0] || h->step_graph[0][1] || h->step_graph[1][0] || h->step_graph[1][1]) {
    if (h->step_graph_stage != h->step_inbox[0] || h->step_graph_stride != stride || h->step_graph_profiling != h->profiling ||
        std::memcmp(&h->step_graph_parameters, p, sizeof(*p)) != 0)
      invalidate_step_graphs(h);
  }
  for (int i = 0; i < g.n_regions; ++i) {   // FastDetector::setThreshold -> std::rint (:21); cv::FAST clamps
    const int t = (int)std::rint(h->thresholds[i]);
    h->h_thr[i] = std::min(std::max(t, 0), 255);
  }
  for (int i = 0; i < 12; ++i) h->h_step_T[i] = T_prior[i];
  const int32_t frame_id = h->step_frame_id = (h->step_frame_id % 0x3fffffff) + 1;   // never 0
  {
    FrameStepState* staged_state = reinterpret_cast<FrameStepState*>(h->h_step_T);   // (its copied prefix)
    staged_state->frame_id = frame_id;
    staged_state->ticket = 0;
    for (int i = 0; i < g.n_regions; ++i) staged_state->thresholds[i] = h->h_thr[i];
  }
  h->sp.localizing = L;
  h->localizing = L;
  static const bool use_graph = std::getenv("VSLAM_NO_FRAME_GRAPH") == nullptr;
  if (use_graph && !h->step_graph[L][buf]) {   // capture the device side of the frame once per frame status and buffer
    const int64_t launches_before = h->launches;
    cudaGraph_t graph = nullptr;
    bool ok = cudaStreamBeginCapture(lane.stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
      h->capturing = true;
      const int issued = issue_frame_step(h, lane, stride, *p, image_left, image_right);
      h->capturing = false;
      ok = cudaStreamEndCapture(lane.stream, &graph) == cudaSuccess && graph != nullptr && issued == VSLAM_OK;
    }
    if (ok) ok = cudaGraphInstantiate(&h->step_graph[L][buf], graph, 0) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    h->step_graph_kernels[L][buf] = (int)(h->launches - launches_before);
    h->launches = launches_before;
    if (!ok) {
      const cudaError_t e = cudaGetLastError();
      h->step_graph[L][buf] = nullptr;
      return fail(VSLAM_ERR_CUDA, "capturing the frame graph failed: %s", cudaGetErrorString(e));
    }
    h->step_graph_stage = h->step_inbox[0];
    h->step_graph_stride = stride;
    h->step_graph_profiling = h->profiling;
    h->step_graph_parameters = *p;
  }
  if (staged) {   // uploaded by vslam_fpg_frame_step_prefetch on the copy stream
    CUDA_TRY(cudaStreamWaitEvent(lane.stream, h->step_inbox_ev[buf], 0));
    --h->step_inbox_count;
  } else {
    if ((rc = upload_pair(h, buf, left, right, image_bytes, lane.stream))) return rc;
  }
  h->step_inbox_head = buf ^ 1;   // the next frame's images go to the other buffer
  if (use_graph) {
    CUDA_TRY(cudaGraphLaunch(h->step_graph[L][buf], lane.stream));
    h->launches += h->step_graph_kernels[L][buf];
    ++h->graph_launches;
  } else if ((rc = issue_frame_step(h, lane, stride, *p, image_left, image_right))) {
    return rc;
  }
  const FrameStepHeader* hd = reinterpret_cast<const FrameStepHeader*>(h->h_step);
  // The last kernel echoes the frame number into the pinned header once every block has published its share
  // (frame_assemble_kernel): polling that word returns a few microseconds before the stream reports idle.  The stream is
  // asked now and then, so an error or a device that never answers still ends in cudaStreamSynchronize.
  bool complete = false;
  if (h->step_poll && !h->profiling) {
    const volatile int32_t* done = &hd->done_frame;
    for (uint32_t spins = 1; !complete; ++spins) {
      complete = *done == frame_id;
      if (!complete && (spins & 0x3ffu) == 0 && cudaStreamQuery(lane.stream) != cudaErrorNotReady) break;
    }
  }
  if (!complete) CUDA_TRY(cudaStreamSynchronize(lane.stream));
  CUDA_TRY(cudaGetLastError());
  collect_clock(h, true, true);
  if (h->profiling) {
    add_interval(h, kKTrack, kEvTrack0, kEvTrack1, 3);
    add_interval(h, kKFrameAligner, kEvTrack1, kEvMatch0, 1);   // (one chain: the aligner sits between track() and the match)
  }
  if ((rc = check_flag(h)) || hd->overflow) {
    cudaMemset(h->b.error_flag, 0, sizeof(int32_t));
    cudaMemset(h->d_step, 0, sizeof(FrameStepState));   // the device-resident points are not usable: a new sequence
    h->step_points = 0;
    h->initialized = false;
    if (rc) return rc;
    return fail(VSLAM_ERR_CAPACITY, "more than %d points in a frame: use the stepwise calls", h->step_cap);
  }
  finish_detection(h, L);
  h->n_device_tracks = hd->n_kept;
  h->step_points = hd->n_points;
  std::memset(out, 0, sizeof(*out));
  out->n_left = h->h_n_desc[0];
  out->n_right = h->h_n_desc[1];
  out->n_previous = hd->n_previous;
  out->n_tracked = hd->stats[0];
  out->n_lost = hd->stats[1];
  out->n_tracked_landmarks = hd->stats[2];
  out->average_descriptor_distance = hd->stats[0] ? (double)hd->stats[3] / (double)hd->stats[0] : std::nan("");
  out->n_tracks = hd->n_kept;
  out->n_new_points = hd->n_out[0];
  out->n_matches = hd->n_out[1];
  out->aligner_rounds = hd->ctl.rounds;
  out->aligner_converged = hd->ctl.converged;
  out->aligner_inliers = (int32_t)std::llrint(hd->system[28]);
  out->aligner_outliers = hd->stats[0] - out->aligner_inliers;
  out->inliers_only = hd->inliers_only;
  out->aligner_total_error = hd->system[27];
  for (int i = 0; i < 12; ++i) out->previous_to_current[i] = hd->ctl.T[i];
  if (hd->ctl.converged) std::memcpy(out->information, hd->ctl.H, sizeof(double) * 36);
  out->tracks = reinterpret_cast<const vslam_track*>(h->h_step + h->step_off_tracks);
  out->kept = h->h_step + h->step_off_kept;
  out->errors = reinterpret_cast<const double*>(h->h_step + h->step_off_errors);
  out->inliers = h->h_step + h->step_off_inliers;
  out->lost = reinterpret_cast<const int32_t*>(h->h_step + h->step_off_lost);
  out->points = reinterpret_cast<const vslam_framepoint*>(h->h_step + h->step_off_points);
  out->frame_points = p->publish_frame_points ? reinterpret_cast<const vslam_previous_point*>(h->h_step + h->step_off_frame_points) : nullptr;
  return VSLAM_OK;
}

int vslam_fpg_recover_points(vslam_fpg* h, const vslam_previous_point* lost, int32_t n_lost, const double W[12],
                             double minimum_depth_meters, double maximum_depth_meters,
                             double maximum_descriptor_distance_tracking, vslam_recovered_point* recovered,
                             int32_t capacity, int32_t* n_recovered) {
  VSLAM_NVTX("vslam_fpg_recover_points [PoseTracker3D::compute->framepoint_generator->recoverPoints]");
  if (!h || !W || n_lost < 0 || (n_lost > 0 && !lost)) return fail(VSLAM_ERR_INVALID_ARGUMENT, "bad arguments");
  if (!h->initialized || h->last_pairs != 1) return fail(VSLAM_ERR_STATE, "recoverPoints without initialize");
  CUDA_TRY(cudaSetDevice(h->device));
  Lane& lane = h->lanes[0];
  int rc = ensure_previous_capacity(h, n_lost);
  if (rc) return rc;
  if (n_lost)
    CUDA_TRY(cudaMemcpyAsync(h->d_previous, lost, sizeof(PreviousPoint) * (size_t)n_lost, cudaMemcpyHostToDevice, lane.stream));
  RecoverParams rp;
  for (int i = 0; i < 12; ++i) rp.W[i] = W[i];
  rp.min_depth = minimum_depth_meters;
  rp.max_depth = maximum_depth_meters;
  rp.max_distance_tracking = maximum_descriptor_distance_tracking;
  rp.brief_tests = h->d_brief_tests;
  // the blurred images of the last initialize() are still in lane 0's scratch (pair 0)
  launch_recover(h->g, h->sp, h->b, 0, lane.blurred, h->d_previous, n_lost, rp, h->d_recover_xy, h->d_recover_n,
                 h->d_recover_desc, h->d_recovered, h->d_recover_n + 2, lane.stream);
  h->launches += n_lost > 0 ? 3 : 1;
  int32_t n = 0;
  CUDA_TRY(cudaMemcpyAsync(&n, h->d_recover_n + 2, sizeof(int32_t), cudaMemcpyDeviceToHost, lane.stream));
  CUDA_TRY(cudaStreamSynchronize(lane.stream));
  CUDA_TRY(cudaGetLastError());
  if (n_recovered) *n_recovered = n;
  if (n > capacity) return fail(VSLAM_ERR_CAPACITY, "capacity %d < %d recovered points", capacity, n);
  if (n && !recovered) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null output");
  if (n) CUDA_TRY(cudaMemcpy(recovered, h->d_recovered, sizeof(RecoveredRecord) * (size_t)n, cudaMemcpyDeviceToHost));
  return VSLAM_OK;
}

int vslam_fpg_set_remaining_features(vslam_fpg* h, int side, const vslam_keypoint* remaining, int32_t n) {
  if (!h || (n > 0 && !remaining) || n < 0) return fail(VSLAM_ERR_INVALID_ARGUMENT, "bad arguments");
  if (side != 0 && side != 1) return fail(VSLAM_ERR_INVALID_ARGUMENT, "side must be 0 or 1");
  if (!h->initialized || h->last_pairs != 1) return fail(VSLAM_ERR_STATE, "set_remaining_features without initialize");
  CUDA_TRY(cudaSetDevice(h->device));
  const int total = h->h_n_desc[side];
  std::vector<uint32_t> xy;
  int rc = fetch_xy(h, side, total, xy);
  if (rc) return rc;
  std::vector<uint8_t> flags(std::max(total, 1), 1);   // 1 = pruned
  for (int i = 0; i < n; ++i) {
    const uint32_t key = ((uint32_t)(int)remaining[i].y << 16) | (uint32_t)(int)remaining[i].x;   // row/col = trunc(pt)
    const auto it = std::lower_bound(xy.begin(), xy.end(), key);   // the device order is ascending (row, col)
    if (it == xy.end() || *it != key)
      return fail(VSLAM_ERR_INVALID_ARGUMENT, "remaining feature (%g, %g) is not a feature of this frame", remaining[i].x, remaining[i].y);
    flags[it - xy.begin()] = 0;
  }
  uint8_t* dst = side == 0 ? h->b.pruned_l : h->b.consumed_r;
  if (total) CUDA_TRY(cudaMemcpy(dst, flags.data(), total, cudaMemcpyHostToDevice));
  return VSLAM_OK;
}

int vslam_fpg_reset_features(vslam_fpg* h) {
  if (!h) return fail(VSLAM_ERR_INVALID_ARGUMENT, "StereoFramePointGenerator::initialize|called with empty frame");
  if (!h->initialized || h->last_pairs != 1) return fail(VSLAM_ERR_STATE, "reset_features without initialize");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = h->lanes[0].stream;
  CUDA_TRY(cudaMemsetAsync(h->b.pruned_l, 0, h->g.cap, s));
  CUDA_TRY(cudaMemsetAsync(h->b.consumed_r, 0, h->g.cap, s));
  h->n_device_tracks = -1;   // the tracks of the abandoned attempt must not pre-load the next compute()
  return VSLAM_OK;
}

int vslam_fpg_get_matches(vslam_fpg* h, vslam_framepoint* out, int32_t capacity, int32_t* n_out) {
  if (!h) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null handle");
  if (!h->initialized || h->last_pairs != 1) return fail(VSLAM_ERR_STATE, "no single-pair compute to read");
  CUDA_TRY(cudaSetDevice(h->device));
  Lane& lane = h->lanes[0];
  launch_emit_matches(h->g, h->sp, h->b, 0, h->n_passes, h->d_matches, h->g.cap, h->d_n_matches, lane.stream);
  ++h->launches;
  int32_t n = 0;
  CUDA_TRY(cudaMemcpyAsync(&n, h->d_n_matches, sizeof(int32_t), cudaMemcpyDeviceToHost, lane.stream));
  CUDA_TRY(cudaStreamSynchronize(lane.stream));
  if (n_out) *n_out = n;
  if (n > capacity) return fail(VSLAM_ERR_CAPACITY, "capacity %d < %d matches", capacity, n);
  if (n) CUDA_TRY(cudaMemcpy(out, h->d_matches, sizeof(FramePointRecord) * n, cudaMemcpyDeviceToHost));
  return remap_records(h, 0, out, n);
}

int vslam_fpg_set_profiling(vslam_fpg* h, int enabled) {
  if (!h) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null handle");
  h->profiling = enabled != 0;
  if (enabled < 0) {   // reset the accumulators
    h->t_detect = h->t_describe = h->t_match = 0;
    for (int i = 0; i < kNumKernels; ++i) h->k_ms[i] = 0, h->k_n[i] = 0;
  }
  return VSLAM_OK;
}

int vslam_fpg_get_time_consumption(vslam_fpg* h, double* detect, double* describe, double* match) {
  if (!h) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null handle");
  if (detect) *detect = h->t_detect;
  if (describe) *describe = h->t_describe;
  if (match) *match = h->t_match;
  return VSLAM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// batched form
// ---------------------------------------------------------------------------------------------------------------

static int check_batch(vslam_fpg* h, int32_t n_pairs) {
  if (!h) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null handle");
  if (n_pairs < 1 || n_pairs > h->max_batch) return fail(VSLAM_ERR_INVALID_ARGUMENT, "n_pairs %d outside [1, max_batch=%d]", n_pairs, h->max_batch);
  return VSLAM_OK;
}

// lanes alternate over chunks; before a lane is reused its previous work is already ordered on its stream
static int batch_pipeline(vslam_fpg* h, int32_t n_pairs, const uint8_t* left, const uint8_t* right, size_t stride,
                          size_t pair_stride, bool upload, bool run, vslam_framepoint* out, int32_t out_capacity) {
  int li = 0;
  int frc = fork_lanes(h);
  if (frc) return frc;
  for (int p0 = 0; p0 < n_pairs; p0 += h->chunk, li = (li + 1) % kLanes) {
    const int n = std::min(h->chunk, n_pairs - p0);
    Lane& lane = h->lanes[h->profiling ? 0 : li];
    if (upload) {
      int rc = upload_images(h, lane, p0, n, left, right, stride, pair_stride);
      if (rc) return rc;
    }
    if (run) {
      run_detect_describe(h, lane, p0, n);
      run_match_select(h, lane, p0, n, nullptr, 0);
      if (h->profiling) {
        CUDA_TRY(cudaStreamSynchronize(lane.stream));
        collect_clock(h, true, true);
      }
    }
    if (out) {
      if (out_capacity == h->out_cap) {
        CUDA_TRY(cudaMemcpyAsync(out + (size_t)p0 * out_capacity, h->d_out + (size_t)p0 * h->out_cap,
                                 sizeof(FramePointRecord) * (size_t)n * h->out_cap, cudaMemcpyDeviceToHost, lane.stream));
      } else {
        const size_t w = sizeof(FramePointRecord) * (size_t)std::min(out_capacity, h->out_cap);
        CUDA_TRY(cudaMemcpy2DAsync(out + (size_t)p0 * out_capacity, sizeof(FramePointRecord) * (size_t)out_capacity,
                                   h->d_out + (size_t)p0 * h->out_cap, sizeof(FramePointRecord) * (size_t)h->out_cap, w, n,
                                   cudaMemcpyDeviceToHost, lane.stream));
      }
    }
  }
  return join_lanes(h);
}

static int batch_finish(vslam_fpg* h, int32_t n_pairs) {
  // counts travel on lane 0, which batch_pipeline ordered after every other lane
  cudaStream_t s = h->lanes[0].stream;
  CUDA_TRY(cudaMemcpyAsync(h->h_n_out, h->b.n_out, sizeof(int32_t) * 2 * n_pairs, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(h->h_n_desc, h->b.n_desc, sizeof(int32_t) * 2 * n_pairs, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(h->h_counts, h->b.raw_count, sizeof(int32_t) * 2 * n_pairs * h->g.n_regions, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(h->h_flag, h->b.error_flag, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  CUDA_TRY(cudaGetLastError());
  int rc = check_flag(h);
  if (rc) cudaMemset(h->b.error_flag, 0, sizeof(int32_t));
  return rc;
}

int vslam_fpg_batch_upload(vslam_fpg* h, int32_t n_pairs, const uint8_t* left, const uint8_t* right, size_t stride,
                           size_t pair_stride) {
  VSLAM_NVTX("vslam_fpg_batch_upload");
  int rc = check_batch(h, n_pairs);
  if (rc) return rc;
  if (!left || !right || stride < (size_t)h->g.cols) return fail(VSLAM_ERR_INVALID_ARGUMENT, "bad image arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  rc = batch_pipeline(h, n_pairs, left, right, stride, pair_stride, true, false, nullptr, 0);
  if (rc) return rc;
  CUDA_TRY(cudaStreamSynchronize(h->lanes[0].stream));   // the host buffers may be reused after this call
  return VSLAM_OK;
}

int vslam_fpg_batch_run(vslam_fpg* h, int32_t n_pairs, int localizing) {
  VSLAM_NVTX("vslam_fpg_batch_run [KeypointDetection + DescriptorExtraction + StereoMatching]");
  int rc = check_batch(h, n_pairs);
  if (rc) return rc;
  if (!h->g.enable_binning) return fail(VSLAM_ERR_INVALID_ARGUMENT, "batched runs require enable_keypoint_binning");
  CUDA_TRY(cudaSetDevice(h->device));
  h->localizing = localizing != 0;
  h->sp.localizing = h->localizing;
  rc = batch_pipeline(h, n_pairs, nullptr, nullptr, 0, 0, false, true, nullptr, 0);
  if (rc) return rc;
  h->initialized = true;
  h->feat_valid = false;
  h->last_pairs = n_pairs;
  return VSLAM_OK;
}

int vslam_fpg_batch_download(vslam_fpg* h, int32_t n_pairs, vslam_framepoint* out, int32_t capacity_per_pair,
                             int32_t* n_framepoints, int32_t* n_matches, int32_t* n_left, int32_t* n_right) {
  VSLAM_NVTX("vslam_fpg_batch_download");
  int rc = check_batch(h, n_pairs);
  if (rc) return rc;
  if (!h->initialized || h->last_pairs < n_pairs) return fail(VSLAM_ERR_STATE, "no batched run to download");
  CUDA_TRY(cudaSetDevice(h->device));
  if (out) {
    rc = batch_pipeline(h, n_pairs, nullptr, nullptr, 0, 0, false, false, out, capacity_per_pair);
    if (rc) return rc;
  }
  if ((rc = batch_finish(h, n_pairs))) return rc;
  for (int i = 0; i < n_pairs; ++i) {
    if (n_framepoints) n_framepoints[i] = h->h_n_out[2 * i];
    if (n_matches) n_matches[i] = h->h_n_out[2 * i + 1];
    if (n_left) n_left[i] = h->h_n_desc[2 * i];
    if (n_right) n_right[i] = h->h_n_desc[2 * i + 1];
    if (out && h->h_n_out[2 * i] > capacity_per_pair)
      return fail(VSLAM_ERR_CAPACITY, "capacity_per_pair %d < %d framepoints of pair %d", capacity_per_pair, h->h_n_out[2 * i], i);
    if (out && (rc = remap_records(h, i, out + (size_t)i * capacity_per_pair, h->h_n_out[2 * i]))) return rc;
  }
  return VSLAM_OK;
}

int vslam_fpg_batch_process(vslam_fpg* h, int32_t n_pairs, const uint8_t* left, const uint8_t* right, size_t stride,
                            size_t pair_stride, int localizing, vslam_framepoint* out, int32_t capacity_per_pair,
                            int32_t* n_framepoints) {
  VSLAM_NVTX("vslam_fpg_batch_process [host images -> framepoints]");
  int rc = check_batch(h, n_pairs);
  if (rc) return rc;
  if (!left || !right || stride < (size_t)h->g.cols) return fail(VSLAM_ERR_INVALID_ARGUMENT, "bad image arguments");
  if (!h->g.enable_binning) return fail(VSLAM_ERR_INVALID_ARGUMENT, "batched runs require enable_keypoint_binning");
  CUDA_TRY(cudaSetDevice(h->device));
  h->localizing = localizing != 0;
  h->sp.localizing = h->localizing;
  rc = batch_pipeline(h, n_pairs, left, right, stride, pair_stride, true, true, out, capacity_per_pair);
  if (rc) return rc;
  h->initialized = true;
  h->feat_valid = false;
  h->last_pairs = n_pairs;
  if ((rc = batch_finish(h, n_pairs))) return rc;
  for (int i = 0; i < n_pairs; ++i) {
    if (n_framepoints) n_framepoints[i] = h->h_n_out[2 * i];
    if (out && h->h_n_out[2 * i] > capacity_per_pair)
      return fail(VSLAM_ERR_CAPACITY, "capacity_per_pair %d < %d framepoints of pair %d", capacity_per_pair, h->h_n_out[2 * i], i);
    if (out && (rc = remap_records(h, i, out + (size_t)i * capacity_per_pair, h->h_n_out[2 * i]))) return rc;
  }
  return VSLAM_OK;
}

int vslam_fpg_batch_linearize(vslam_fpg* h, int32_t n_pairs, const double T[12], int ignore_outliers,
                              double maximum_error_kernel, double minimum_reliable_depth, double maximum_reliable_depth,
                              int enable_inverse_depth_as_information, int32_t rounds) {
  VSLAM_NVTX("vslam_fpg_batch_linearize [PoseTracker3D::compute->pose_optim]");
  int rc = check_batch(h, n_pairs);
  if (rc) return rc;
  if (!T || rounds < 1) return fail(VSLAM_ERR_INVALID_ARGUMENT, "bad pose / rounds");
  if (!h->initialized || h->last_pairs < n_pairs) return fail(VSLAM_ERR_STATE, "no batched run to align");
  CUDA_TRY(cudaSetDevice(h->device));
  AlignerCamera cam;   // StereoUVAligner::initialize :65-68
  const double K[9] = {h->cfg.fx, 0, h->cfg.cx, 0, h->cfg.fy, h->cfg.cy, 0, 0, 1};
  for (int i = 0; i < 9; ++i) cam.K[i] = K[i];
  cam.baseline[0] = h->cfg.bx;
  cam.baseline[1] = cam.baseline[2] = 0;
  cam.rows = h->g.rows;
  cam.cols = h->g.cols;
  cam.min_depth = minimum_reliable_depth;
  Lane& lane = h->lanes[0];
  mark(h, lane, kEvLin0);
  for (int r = 0; r < rounds; ++r) {
    launch_linearize_pairs(h->d_out, h->out_cap, h->b.n_out, n_pairs, cam, T, ignore_outliers, maximum_error_kernel,
                           maximum_reliable_depth, enable_inverse_depth_as_information, h->d_systems, h->d_pair_errors,
                           h->d_pair_inliers, lane.stream);
    ++h->launches;
  }
  mark(h, lane, kEvLin1);
  CUDA_TRY(cudaGetLastError());
  if (h->profiling) {
    CUDA_TRY(cudaStreamSynchronize(lane.stream));
    add_interval(h, kKLinearize, kEvLin0, kEvLin1, rounds);
  }
  return VSLAM_OK;
}

int vslam_fpg_batch_get_systems(vslam_fpg* h, int32_t n_pairs, vslam_linear_system* systems, double* errors,
                                uint8_t* inliers, int32_t capacity_per_pair) {
  int rc = check_batch(h, n_pairs);
  if (rc) return rc;
  if (!systems) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null output");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = h->lanes[0].stream;
  CUDA_TRY(cudaMemcpyAsync(h->h_systems, h->d_systems, sizeof(double) * 32 * n_pairs, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(h->h_n_out, h->b.n_out, sizeof(int32_t) * 2 * n_pairs, cudaMemcpyDeviceToHost, s));
  if (errors || inliers) {
    const size_t w = (size_t)std::min(capacity_per_pair, h->out_cap);
    if (errors)
      CUDA_TRY(cudaMemcpy2DAsync(errors, sizeof(double) * capacity_per_pair, h->d_pair_errors, sizeof(double) * h->out_cap,
                                 sizeof(double) * w, n_pairs, cudaMemcpyDeviceToHost, s));
    if (inliers)
      CUDA_TRY(cudaMemcpy2DAsync(inliers, capacity_per_pair, h->d_pair_inliers, h->out_cap, w, n_pairs,
                                 cudaMemcpyDeviceToHost, s));
  }
  CUDA_TRY(cudaStreamSynchronize(s));
  for (int p = 0; p < n_pairs; ++p) {
    const double* v = h->h_systems + (size_t)p * 32;
    vslam_linear_system& o = systems[p];
    int k = 0;
    for (int i = 0; i < 6; ++i)
      for (int j = i; j < 6; ++j, ++k) o.H[i * 6 + j] = o.H[j * 6 + i] = v[k];
    for (int i = 0; i < 6; ++i) o.b[i] = v[21 + i];
    o.total_error = v[27];
    o.number_of_inliers = (int32_t)std::llrint(v[28]);
    o.number_of_outliers = h->h_n_out[2 * p] - o.number_of_inliers;
  }
  return VSLAM_OK;
}

int vslam_fpg_get_kernel_profile(vslam_fpg* h, double* milliseconds, int64_t* launches) {
  if (!h) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null handle");
  for (int i = 0; i < kNumKernels; ++i) {
    if (milliseconds) milliseconds[i] = h->k_ms[i];
    if (launches) launches[i] = h->k_n[i];
  }
  return VSLAM_OK;
}

void* vslam_fpg_stream(vslam_fpg* h) { return h ? (void*)h->lanes[0].stream : nullptr; }

int vslam_fpg_synchronize(vslam_fpg* h) {
  if (!h) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  for (auto& l : h->lanes) CUDA_TRY(cudaStreamSynchronize(l.stream));
  return VSLAM_OK;
}

int64_t vslam_fpg_launch_count(const vslam_fpg* h) { return h ? h->launches : 0; }

int vslam_fpg_debug_keypoint_mask(vslam_fpg* h, int32_t pair, int side, uint32_t* words) {
  if (!h || !words) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  if (!h->initialized || pair < 0 || pair >= h->last_pairs || h->last_pairs > h->chunk)
    return fail(VSLAM_ERR_STATE, "debug taps need a run of at most chunk=%d pairs", h->chunk);
  CUDA_TRY(cudaSetDevice(h->device));
  const Geometry& g = h->g;
  const int w = (g.cols + 31) / 32;
  CUDA_TRY(cudaMemcpy2D(words, sizeof(uint32_t) * w, h->lanes[0].mask + ((size_t)2 * pair + side) * g.rows * g.mask_words,
                        sizeof(uint32_t) * g.mask_words, sizeof(uint32_t) * w, g.rows, cudaMemcpyDeviceToHost));
  return VSLAM_OK;
}

int vslam_fpg_debug_blurred(vslam_fpg* h, int32_t pair, int side, uint8_t* image) {
  if (!h || !image) return fail(VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  if (!h->initialized || pair < 0 || pair >= h->last_pairs || h->last_pairs > std::min(h->chunk, h->describe_sub_chunk))
    return fail(VSLAM_ERR_STATE, "the blurred-image tap needs a run of at most %d pairs", std::min(h->chunk, h->describe_sub_chunk));
  if (h->d_brief_tests) return fail(VSLAM_ERR_STATE, "no blurred image with the BRIEF-32 extractor");
  CUDA_TRY(cudaSetDevice(h->device));
  const Geometry& g = h->g;
  CUDA_TRY(cudaMemcpy2D(image, g.cols, h->lanes[0].blurred + ((size_t)2 * pair + side) * g.rows * g.pitch, g.pitch, g.cols,
                        g.rows, cudaMemcpyDeviceToHost));
  return VSLAM_OK;
}

double vslam_threshold_proposal(double threshold, int32_t n_keypoints, double target, double tolerance,
                                double maximum_change, double threshold_minimum, double threshold_maximum) {
  return threshold_proposal(threshold, n_keypoints, target, tolerance, maximum_change, threshold_minimum, threshold_maximum);
}

}  // extern "C"
