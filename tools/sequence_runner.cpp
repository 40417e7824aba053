// sequence_runner.cpp -- the tracker's per-frame order (reference src/position_tracking/pose_tracker_3d.cpp:80, 239,
// 124-126 / 355-357, 437-472, 210: initialize -> track against ALL points of the previous frame -> StereoUVAligner
// initialize + converge over the tracks -> _prunePoints on the device records -> compute with the surviving tracks
// pre-loaded) driven from C++14 through include/vslam_b200.hpp, i.e. what the adapters do per frame minus the reference's
// object graph.  bench.py builds and runs it to report the single-sequence latency without the Python harness in the loop.
//
//   sequence_runner <frames.u8> <n_frames> <warmup> <24 configuration numbers, see below> [passes [fused]]
// fused = 1: the same per-frame order as ONE device pass per frame (vslam_fpg_frame_step through
// StereoFramePointGenerator::trackFrame: one graph launch and one synchronisation, points() resident on the device).
// fused = 2: the same, and the images of frame k + 1 are uploaded while frame k runs (prefetchFrame: a replayed sequence
// knows its next frame, as the reference's playback from disk does).
// passes > 1 replays the sequence from its first frame (fresh tracker state) so that a short sequence gives a timed
// region long enough to overlap with the other GPUs' runs (BASELINE configs[4]: one sequence per GPU).
// frames.u8: [n_frames][2][rows][cols] u8.  Prints one JSON object.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include "vslam_b200.hpp"

int main(int argc, char** argv) {
  if (argc != 4 + 24 && argc != 4 + 25 && argc != 4 + 26) {
    std::fprintf(stderr, "usage: sequence_runner frames.u8 n_frames warmup rows cols tolerance thr_min thr_max max_change "
                         "detectors_v detectors_h binning bin_size max_distance min_disparity max_offset fx fy cx cy bx "
                         "projection_tracking_distance error_delta_for_convergence maximum_error_kernel damping "
                         "minimum_number_of_inliers maximum_reliable_depth_meters\n");
    return 2;
  }
  try {
    const int n_frames = std::atoi(argv[2]), warmup = std::atoi(argv[3]);
    char** a = argv + 4;
    vslam_fpg_config c = {};
    c.rows = std::atoi(a[0]); c.cols = std::atoi(a[1]);
    c.target_number_of_keypoints_tolerance = std::atof(a[2]);
    c.detector_threshold_minimum = std::atoi(a[3]); c.detector_threshold_maximum = std::atoi(a[4]);
    c.detector_threshold_maximum_change = std::atof(a[5]);
    c.number_of_detectors_vertical = std::atoi(a[6]); c.number_of_detectors_horizontal = std::atoi(a[7]);
    c.enable_keypoint_binning = std::atoi(a[8]); c.bin_size_pixels = std::atoi(a[9]);
    c.maximum_matching_distance_triangulation = std::atof(a[10]); c.minimum_disparity_pixels = std::atof(a[11]);
    c.maximum_epipolar_search_offset_pixels = std::atoi(a[12]);
    c.fx = std::atof(a[13]); c.fy = std::atof(a[14]); c.cx = std::atof(a[15]); c.cy = std::atof(a[16]); c.bx = std::atof(a[17]);
    const int tracking_distance = std::atoi(a[18]);
    const double descriptor_distance = 25.6;
    vslam_aligner_parameters ap = {};
    ap.error_delta_for_convergence = std::atof(a[19]);
    ap.maximum_error_kernel = std::atof(a[20]);
    ap.damping = std::atof(a[21]);
    ap.maximum_number_of_iterations = 1000;
    ap.minimum_number_of_inliers = std::atoi(a[22]);
    const double maximum_reliable_depth = std::atof(a[23]);
    const int passes = argc >= 4 + 25 ? std::atoi(a[24]) : 1;
    const int fused_mode = argc == 4 + 26 ? std::atoi(a[25]) : 0;
    const bool fused = fused_mode != 0, prefetch = fused_mode == 2;
    vslam_frame_step_parameters fsp = {};
    fsp.track_by_appearance = 0;
    fsp.projection_tracking_distance_pixels = tracking_distance;
    fsp.maximum_descriptor_distance_tracking = descriptor_distance;
    fsp.aligner = ap;
    fsp.enable_inverse_depth_as_information = 1;
    fsp.minimum_track_length_for_landmark_creation = 1;
    fsp.maximum_reliable_depth_meters = maximum_reliable_depth;
    fsp.minimum_reliable_depth_meters = 0.1;
    // points() of the frame as 128-byte records for a host that mirrors them (the device keeps its own copy either way)
    fsp.publish_frame_points = std::getenv("VSLAM_RUNNER_NO_FRAME_POINTS") ? 0 : 1;
    const size_t image_bytes = (size_t)c.rows * c.cols;

    // page-locked frame buffers (vslam_host_alloc): the H2D copy of initialize() is one asynchronous DMA
    uint8_t* frames = nullptr;
    if (vslam_host_alloc(reinterpret_cast<void**>(&frames), 2 * image_bytes * n_frames) != VSLAM_OK)
      throw std::runtime_error(vslam_last_error());
    {
      std::ifstream f(argv[1], std::ios::binary);
      if (!f.read(reinterpret_cast<char*>(frames), (std::streamsize)(2 * image_bytes * n_frames)))
        throw std::runtime_error("short frames file");
    }
    vslam::StereoFramePointGenerator generator(c);
    generator.setProjectionTrackingDistancePixels(tracking_distance);
    generator.setMaximumDescriptorDistanceTracking(descriptor_distance);
    // the band world moves the camera a quarter baseline per frame along x (tests/cpp/host_api_check.cpp)
    const double tx = -(-c.bx / c.fx) / 4;
    const std::array<double, 12> motion{{1, 0, 0, tx, 0, 1, 0, 0, 0, 0, 1, 0}};

    // StereoUVAligner wired as slam_assembly.cpp:68-71: minimum reliable depth = the generator's minimum_depth_meters
    vslam::StereoUVAligner aligner(ap, 1 << 16);
    const double K[9] = {c.fx, 0, c.cx, 0, c.fy, c.cy, 0, 0, 1};
    const double baseline[3] = {c.bx, 0, 0};
    std::vector<double> moving, fixed, omega, weights;
    double s_align = 0;
    long n_rounds = 0, n_inliers = 0;
    double worst_translation_error = 0;

    vslam::Frame previous, current;
    bool have_previous = false;
    double seconds = 0, s_initialize = 0, s_track = 0, s_compute = 0, s_assemble = 0;
    long n_previous = 0, n_tracks = 0, n_new = 0;
    auto since = [](std::chrono::steady_clock::time_point t) {
      return std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count();
    };
    std::vector<int32_t> lost;
    // VSLAM_RUNNER_PROFILE=1: device time per stage of the fused frame (CUDA events between the kernels; the frame then
    // runs as ONE chain without parallel branches, so its total is not the number to quote)
    const bool stage_profile = fused && std::getenv("VSLAM_RUNNER_PROFILE") != nullptr;
    if (stage_profile) vslam_fpg_set_profiling(generator.handle(), 1);
    for (int pass = 0; pass < passes; ++pass) {
    have_previous = false;
    if (fused) generator.resetSequence();
    if (prefetch) generator.prefetchFrame(frames, frames + image_bytes, (size_t)c.cols);
    for (int k = 0; k < n_frames; ++k) {
      const auto t0 = std::chrono::steady_clock::now();
      current.status = k == 0 ? vslam::Frame::Localizing : vslam::Frame::Tracking;
      current.intensity_image_left = frames + (size_t)(2 * k) * image_bytes;
      current.intensity_image_right = frames + (size_t)(2 * k + 1) * image_bytes;
      current.image_step = (size_t)c.cols;
      current.tracks.clear();
      if (fused) {
        if (prefetch) {   // frame k was staged by the previous iteration; frame k + 1 travels while k runs
          if (k + 1 < n_frames)
            generator.prefetchFrame(frames + (size_t)(2 * k + 2) * image_bytes, frames + (size_t)(2 * k + 3) * image_bytes,
                                    (size_t)c.cols);
          current.intensity_image_left = current.intensity_image_right = nullptr;
        }
        // the whole frame on the device; trackFrame copies tracks / points / points() of the frame out of the pinned block
        const vslam_frame_step_result r = generator.trackFrame(&current, motion, fsp);
        const auto t1 = std::chrono::steady_clock::now();
        if (k >= warmup || pass > 0) {
          seconds += std::chrono::duration<double>(t1 - t0).count();
          n_rounds += r.aligner_rounds;
          n_inliers += r.aligner_inliers;
          n_previous += r.n_previous;
          n_tracks += r.n_tracks;
          n_new += r.n_new_points;
          const double e = std::fabs(r.previous_to_current[3] - tx);
          if (r.n_tracked && e > worst_translation_error) worst_translation_error = e;
        }
        continue;
      }
      generator.initialize(&current);
      const double d_initialize = since(t0);
      const auto t_track = std::chrono::steady_clock::now();
      if (have_previous) generator.track(&current, &previous, motion, lost, false);
      const double d_track = since(t_track);
      // pose optimisation over the tracks (pose_tracker_3d.cpp:355-357): StereoUVAligner::initialize packs per point
      // (stereouv_aligner.cpp:26-64), converge() runs fused on the device, then errors / inliers as _prunePoints reads them
      const auto t_align = std::chrono::steady_clock::now();
      int rounds = 0, inliers = 0;
      if (have_previous && !current.tracks.empty()) {
        const size_t n = current.tracks.size();
        moving.resize(3 * n); fixed.resize(4 * n); omega.assign(n, 1.0); weights.resize(n);
        for (size_t u = 0; u < n; ++u) {
          const vslam_track& t = current.tracks[u];
          const vslam_previous_point& q = previous.previous_points[t.index_previous];
          for (int d = 0; d < 3; ++d) moving[3 * u + d] = q.camera_left[d];
          fixed[4 * u] = t.xl; fixed[4 * u + 1] = t.yl; fixed[4 * u + 2] = t.xr; fixed[4 * u + 3] = t.yr;
          const double w = maximum_reliable_depth / t.camera[2];
          weights[u] = w < 1.0 ? w : 1.0;
        }
        aligner.initialize((int32_t)n, moving.data(), fixed.data(), omega.data(), weights.data(), K, baseline, c.rows,
                           c.cols, 0.1, motion);
        aligner.converge();
        // _prunePoints (:437-472): rejected tracks leave the frame and, on the device, the bin pre-load of compute()
        generator.pruneTracks(&current, aligner.handle(), ap.maximum_error_kernel);
        rounds = aligner.numberOfRounds();
        inliers = aligner.numberOfInliers();
        const double e = std::fabs(aligner.previousToCurrent()[3] - tx);
        if ((k >= warmup || pass > 0) && e > worst_translation_error) worst_translation_error = e;
      }
      const double d_align = since(t_align);
      const auto t_compute = std::chrono::steady_clock::now();
      generator.compute(&current);
      const double d_compute = since(t_compute);
      const auto t_assemble = std::chrono::steady_clock::now();
      // points() of this frame as the next frame's track() reads them: the tracks, then the new points
      // (what GpuStereoFramePointGenerator::fillPreviousPoint does per FramePoint)
      current.previous_points.resize(current.tracks.size() + current.points.size());
      size_t i = 0;
      auto fill = [&](const double camera[3], int32_t index_left, int32_t index_right, int32_t epipolar_offset) {
        vslam_previous_point& q = current.previous_points[i++];
        for (int d = 0; d < 3; ++d) q.camera_left[d] = q.world[d] = camera[d];
        std::memcpy(q.descriptor_left, &current.descriptors_left[(size_t)index_left * VSLAM_DESCRIPTOR_BYTES], VSLAM_DESCRIPTOR_BYTES);
        std::memcpy(q.descriptor_right, &current.descriptors_right[(size_t)index_right * VSLAM_DESCRIPTOR_BYTES], VSLAM_DESCRIPTOR_BYTES);
        q.epipolar_offset = epipolar_offset;
        q.has_landmark = 1;
        q.keypoint_size = 7.f;
        q.reserved = 0;
      };
      for (const vslam_track& t : current.tracks) fill(t.camera, t.index_left, t.index_right, t.epipolar_offset);
      for (const vslam_framepoint& p : current.points) fill(p.camera, p.index_left, p.index_right, p.epipolar_offset);
      const auto t1 = std::chrono::steady_clock::now();
      if (k >= warmup || pass > 0) {
        seconds += std::chrono::duration<double>(t1 - t0).count();
        s_initialize += d_initialize;
        s_track += d_track;
        s_align += d_align;
        n_rounds += rounds;
        n_inliers += inliers;
        s_compute += d_compute;
        s_assemble += since(t_assemble);
        n_previous += have_previous ? (long)previous.previous_points.size() : 0;
        n_tracks += (long)current.tracks.size();
        n_new += (long)current.points.size();
      }
      std::swap(previous, current);
      have_previous = true;
    }
    }
    const int timed = n_frames * passes - warmup;
    if (stage_profile) {
      double ms[VSLAM_FPG_KERNELS];
      int64_t launches[VSLAM_FPG_KERNELS];
      vslam_fpg_get_kernel_profile(generator.handle(), ms, launches);
      const char* names[VSLAM_FPG_KERNELS] = {"fast_nms", "compact", "blur", "describe", "match", "select", "linearize_pairs",
                                              "track", "frame_aligner"};
      const long frames_total = (long)passes * n_frames;
      std::fprintf(stderr, "device us per frame (%ld frames, one chain):", frames_total);
      for (int i = 0; i < VSLAM_FPG_KERNELS; ++i)
        if (launches[i]) std::fprintf(stderr, " %s %.1f", names[i], ms[i] * 1000.0 / frames_total);
      std::fprintf(stderr, "\n");
    }
    std::printf("{\"frames_per_s\": %.3f, \"ms_per_frame\": %.6f, \"mean_previous_points\": %.2f, \"mean_tracks\": %.2f, "
                "\"mean_new_points\": %.2f, \"frames\": %d, \"us_initialize_with_feature_download\": %.1f, \"us_track\": %.1f, "
                "\"us_align\": %.1f, \"us_compute\": %.1f, \"us_assemble_previous_points\": %.1f, \"mean_aligner_rounds\": %.2f, "
                "\"mean_aligner_inliers\": %.1f, \"worst_translation_error_m\": %.3g, \"fused\": %d}\n",
                timed / seconds, seconds / timed * 1e3, (double)n_previous / timed, (double)n_tracks / timed,
                (double)n_new / timed, timed, s_initialize / timed * 1e6, s_track / timed * 1e6, s_align / timed * 1e6,
                s_compute / timed * 1e6, s_assemble / timed * 1e6, (double)n_rounds / timed, (double)n_inliers / timed,
                worst_translation_error, fused_mode);
    vslam_host_free(frames);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "FAILED: %s\n", e.what());
    return 1;
  }
  return 0;
}
