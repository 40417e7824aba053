// frame_step.cu -- the glue kernels of the fused tracked frame (vslam_fpg_frame_step): what PoseTracker3D::compute does
// on the host between the generator's and the aligner's calls (reference src/position_tracking/pose_tracker_3d.cpp:
// 124-126 / 355-357 aligner initialize + converge, 437-472 _prunePoints, 210 compute, and the points() of the frame the
// next track() reads), moved to the device so that a frame is ONE graph launch and one synchronisation.  Counts travel
// through FrameStepState (device memory), never through kernel arguments.
//
//   (StereoUVAligner::initialize over the tracks is the head of converge_cluster_kernel, _prunePoints on the bin
//    pre-load records its tail: aligner.cu, FrameFill / FramePrune)
//   frame_assemble_kernel       points() of the frame (surviving tracks, then the new framepoints, with their descriptors)
//                               replace the previous points in place; everything the host reads is written into the
//                               handle's pinned, device-mapped result block
#include "kernels.cuh"

namespace vslam {

namespace {

// coalesced 16-byte copies of a contiguous range by one block
__device__ __forceinline__ void copy16(void* dst, const void* src, size_t bytes, int tid, int threads) {
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(dst);
  for (size_t i = tid; i < bytes / 16; i += threads) d[i] = s[i];
}

constexpr int kAssembleThreads = 256;
constexpr int kAssemblePoints = 64;   // points() entries per block: 8 KB of shared memory, written out in 512-byte rows

// `part` (kAssembleAll / kAssembleTracks / kAssembleRest) selects the blocks of this launch: the tracks' blocks need the
// prune only and may run beside the bin selection; they do not read its counts (a frame that overflows is discarded by
// the host together with the device-resident points, so what they wrote then is never used).
__device__ __forceinline__ void frame_assemble_block(const Geometry& g, const FrameStepBuffers& f, const FrameStepParams& p,
                                                     int part) {
  __shared__ __align__(16) PreviousPoint s_points[kAssemblePoints];
  __shared__ __align__(16) TrackRecord s_tracks[kAssemblePoints];
  __shared__ int s_pos[kAssemblePoints];
  const int tid = threadIdx.x;
  const int n_tracks = min(f.stats[0], f.cap);
  const int n_kept = f.state->n_kept;
  const int n_new = part == kAssembleTracks ? 0 : min(f.n_out[0], f.out_cap);
  const bool overflow = f.state->overflow != 0 || n_kept + n_new > f.cap;
  const int n_points = overflow ? 0 : n_kept + n_new;
  const uint8_t* desc_l = f.desc;
  const uint8_t* desc_r = f.desc + (size_t)g.cap * kDescBytes;
  const int track_blocks = (f.cap + kAssemblePoints - 1) / kAssemblePoints;
  const int new_blocks = (f.out_cap + kAssemblePoints - 1) / kAssemblePoints;

  auto fill = [&](PreviousPoint* q, const double camera[3], int index_left, int index_right, int epipolar_offset, int length) {
    for (int d = 0; d < 3; ++d) q->camera[d] = q->world[d] = camera[d];
    const uint4* dl = reinterpret_cast<const uint4*>(desc_l + (size_t)index_left * kDescBytes);
    const uint4* dr = reinterpret_cast<const uint4*>(desc_r + (size_t)index_right * kDescBytes);
    uint4* ql = reinterpret_cast<uint4*>(q->descriptor_left);
    uint4* qr = reinterpret_cast<uint4*>(q->descriptor_right);
    ql[0] = dl[0]; ql[1] = dl[1];
    qr[0] = dr[0]; qr[1] = dr[1];
    q->epipolar_offset = epipolar_offset;
    q->has_landmark = length >= p.min_track_length;
    q->keypoint_size = 7.f;        // cv::FastFeatureDetector keypoints (fast.cpp: KeyPoint(x, y, 7.f, -1, score))
    q->reserved = length;          // trackLength()
  };

  const int block = part == kAssembleRest ? (int)blockIdx.x + track_blocks : (int)blockIdx.x;
  if (block < track_blocks) {
    // ---- tracks [k0, k0 + 64) of track(): the survivors go to their ordered position among points() and the host's tracks.
    // A block's survivors are consecutive positions (the order is kept), so both outputs are contiguous ranges.
    const int k0 = block * kAssemblePoints;
    if (k0 >= n_tracks) return;
    if (tid < kAssemblePoints) {
      const int k = k0 + tid;
      const int pos = k < n_tracks ? f.kept_pos[k] : -1;
      s_pos[tid] = pos;
    }
    __syncthreads();
    // first surviving position of the block and the count
    int first = -1, count = 0;
    for (int i = 0; i < kAssemblePoints; ++i)
      if (s_pos[i] >= 0) {
        if (first < 0) first = s_pos[i];
        ++count;
      }
    if (tid < kAssemblePoints && s_pos[tid] >= 0) {
      const TrackRecord t = f.tracks[k0 + tid];
      const int slot = s_pos[tid] - first;
      s_tracks[slot] = t;
      fill(&s_points[slot], t.camera, t.index_left, t.index_right, t.epipolar_offset, f.track_length[k0 + tid] + 1);
    }
    __syncthreads();
    if (count && !overflow) {
      if (tid < count) f.estimates[first + tid].information_scale = 0.0;   // points() of the new frame: no estimate yet
      copy16(f.previous + first, s_points, sizeof(PreviousPoint) * (size_t)count, tid, kAssembleThreads);
      if (p.publish_frame_points)
        copy16(f.h_frame_points + first, s_points, sizeof(PreviousPoint) * (size_t)count, tid, kAssembleThreads);
      // TrackRecord is 88 bytes: copy as 8-byte words (first * 88 is 8-byte aligned)
      const uint2* s = reinterpret_cast<const uint2*>(s_tracks);
      uint2* d = reinterpret_cast<uint2*>(f.h_tracks + first);
      for (int i = tid; i < count * (int)(sizeof(TrackRecord) / 8); i += kAssembleThreads) d[i] = s[i];
    }
    // per-track results of the aligner and the prune, for the host's own bookkeeping
    if (tid < kAssemblePoints && k0 + tid < n_tracks) {
      const int k = k0 + tid;
      f.h_kept[k] = s_pos[tid] >= 0;
      f.h_errors[k] = f.aligner.errors[k];
      f.h_inliers[k] = f.aligner.inliers[k];
    }
    return;
  }
  const int nb = block - track_blocks;
  if (nb < new_blocks) {
    // ---- new framepoints [k0, k0 + 64) of compute(): behind the surviving tracks
    const int k0 = nb * kAssemblePoints;
    if (k0 >= n_new) return;
    const int count = min(kAssemblePoints, n_new - k0);
    if (tid < count) {
      const FramePointRecord r = f.points[k0 + tid];
      fill(&s_points[tid], r.camera, r.index_left, r.index_right, r.epipolar_offset, 1);
    }
    __syncthreads();
    if (!overflow) {
      if (tid < count) f.estimates[n_kept + k0 + tid].information_scale = 0.0;
      copy16(f.previous + n_kept + k0, s_points, sizeof(PreviousPoint) * (size_t)count, tid, kAssembleThreads);
      if (p.publish_frame_points)
        copy16(f.h_frame_points + n_kept + k0, s_points, sizeof(PreviousPoint) * (size_t)count, tid, kAssembleThreads);
    }
    // FramePointRecord is 56 bytes: 8-byte words
    const uint2* s = reinterpret_cast<const uint2*>(f.points + k0);
    uint2* d = reinterpret_cast<uint2*>(f.h_points + k0);
    for (int i = tid; i < count * (int)(sizeof(FramePointRecord) / 8); i += kAssembleThreads) d[i] = s[i];
    return;
  }
  // ---- last block: lost points, header, and the count the next frame's track() reads
  const int n_lost = min(f.stats[1], f.cap);
  for (int i = tid; i < n_lost; i += kAssembleThreads) f.h_lost[i] = f.lost[i];
  if (tid < 4) f.h_header->stats[tid] = f.stats[tid];
  if (tid < 4) f.h_status[tid] = f.d_status[tid];
  for (int i = tid; i < 2 * f.n_regions; i += kAssembleThreads) f.h_counts[i] = f.d_raw_count[i];
  if (tid < 2) f.h_header->n_out[tid] = f.n_out[tid];
  if (tid < 32) f.h_header->system[tid] = f.aligner.system[tid];
  {
    const double* s = reinterpret_cast<const double*>(f.ctl);
    double* d = reinterpret_cast<double*>(&f.h_header->ctl);
    for (int i = tid; i < (int)(sizeof(GnControl) / 8); i += kAssembleThreads) d[i] = s[i];
  }
  if (tid == 0) {
    f.h_header->n_kept = n_kept;
    f.h_header->inliers_only = f.state->inliers_only;
    f.h_header->overflow = overflow;
    f.h_header->error_flag = *f.error_flag;
    f.h_header->n_previous = f.state->n_previous;
    f.h_header->n_points = n_points;
    f.state->n_previous = n_points;   // (no other block reads it: track() of this frame is long done)
  }
}


// The frame is complete when every block of the LAST kernel has published its share: the blocks take a ticket after a
// device-wide fence (release), the one that draws the last ticket (acquire) issues ONE system-wide fence -- cumulative
// over what it observed -- and echoes the host's frame number into the header.  The host polls that word in pinned memory
// instead of waiting for the stream to drain (vslam_fpg_frame_step).  (A system-wide fence in every block waits for the
// PCIe acknowledgements of that block's own writes: +7 us per frame, measured.)
__global__ void __launch_bounds__(kAssembleThreads) frame_assemble_kernel(Geometry g, FrameStepBuffers f, FrameStepParams p,
                                                                         int part) {
  frame_assemble_block(g, f, p, part);
  if (part == kAssembleTracks) return;   // (joined to the rest of the frame by an edge of the graph)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int ticket = atomicAdd(&f.state->ticket, 1);
    if (ticket == (int)gridDim.x - 1) {
      __threadfence_system();
      *reinterpret_cast<volatile int32_t*>(&f.h_header->done_frame) = f.state->frame_id;
    }
  }
}

}  // namespace

void launch_frame_assemble(const Geometry& g, const FrameStepBuffers& f, const FrameStepParams& p, int part,
                           cudaStream_t stream) {
  const int track_blocks = (f.cap + kAssemblePoints - 1) / kAssemblePoints;
  const int rest_blocks = (f.out_cap + kAssemblePoints - 1) / kAssemblePoints + 1;
  const int blocks = part == kAssembleTracks ? track_blocks : (part == kAssembleRest ? rest_blocks : track_blocks + rest_blocks);
  frame_assemble_kernel<<<blocks, kAssembleThreads, 0, stream>>>(g, f, p, part);
}

}  // namespace vslam
