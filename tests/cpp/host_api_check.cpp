// host_api_check.cpp -- a C++14 consumer of include/vslam_b200.hpp, shaped like the reference's own
// executables/test_stereo_frontend.cpp (initialize -> compute on a stereo pair, prints counts) plus an aligner run.
// Reads raw u8 images and correspondence arrays written by tests/test_gpu_cpp_host.py and prints results as text that
// the pytest compares with the CPU oracle.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>

#include "vslam_b200.hpp"

static std::vector<uint8_t> slurp(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open " + path);
  return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const std::string dir = argv[1];
  try {
    // ---- framepoint generation, configuration_kitti.yaml values (reference configurations/configuration_kitti.yaml:58-93)
    vslam_fpg_config c = {};
    c.rows = 376; c.cols = 1241;
    c.target_number_of_keypoints_tolerance = 0.1;
    c.detector_threshold_minimum = 20; c.detector_threshold_maximum = 100; c.detector_threshold_maximum_change = 0.1;
    c.number_of_detectors_vertical = 1; c.number_of_detectors_horizontal = 1;
    c.enable_keypoint_binning = 1; c.bin_size_pixels = 15;
    c.maximum_matching_distance_triangulation = 51.2; c.minimum_disparity_pixels = 1;
    c.maximum_epipolar_search_offset_pixels = 0;
    c.fx = c.fy = 718.856; c.cx = 607.1928; c.cy = 185.2157; c.bx = -386.1448;
    vslam::StereoFramePointGenerator generator(c);
    const std::vector<uint8_t> left = slurp(dir + "/left.u8"), right = slurp(dir + "/right.u8");
    vslam::Frame frame;
    frame.status = vslam::Frame::Localizing;
    frame.intensity_image_left = left.data();
    frame.intensity_image_right = right.data();
    frame.image_step = 1241;
    generator.initialize(&frame);
    generator.compute(&frame);
    std::printf("features %zu %zu\n", frame.keypoints_left.size(), frame.keypoints_right.size());
    std::printf("points %zu new %d threshold %.1f\n", frame.points.size(), generator.numberOfNewPoints(),
                generator.meanDetectorThreshold());
    unsigned long long h = 1469598103934665603ull;   // FNV-1a over the descriptors and the selected points
    for (uint8_t b : frame.descriptors_left) h = (h ^ b) * 1099511628211ull;
    for (const vslam_framepoint& p : frame.points) {
      const int32_t v[4] = {p.index_left, p.index_right, p.distance, p.epipolar_offset};
      for (int32_t x : v) h = (h ^ (unsigned long long)(uint32_t)x) * 1099511628211ull;
    }
    std::printf("hash %llu\n", h);
    if (!frame.points.empty())
      std::printf("first %.17g %.17g %.17g\n", frame.points[0].camera[0], frame.points[0].camera[1], frame.points[0].camera[2]);

    // ---- next frame: initialize -> track (against the first frame's points) -> compute -> recoverPoints, the order of
    // PoseTracker3D::compute (reference src/position_tracking/pose_tracker_3d.cpp:80, 239, 206-210)
    for (const vslam_framepoint& p : frame.points) {   // frame.points() as the next track() reads them
      vslam_previous_point q = {};
      for (int i = 0; i < 3; ++i) q.camera_left[i] = q.world[i] = p.camera[i];
      for (int i = 0; i < VSLAM_DESCRIPTOR_BYTES; ++i) {
        q.descriptor_left[i] = frame.descriptors_left[(size_t)p.index_left * VSLAM_DESCRIPTOR_BYTES + i];
        q.descriptor_right[i] = frame.descriptors_right[(size_t)p.index_right * VSLAM_DESCRIPTOR_BYTES + i];
      }
      q.epipolar_offset = p.epipolar_offset;
      q.has_landmark = 1;
      q.keypoint_size = 7.f;
      frame.previous_points.push_back(q);
    }
    const std::vector<uint8_t> left1 = slurp(dir + "/left1.u8"), right1 = slurp(dir + "/right1.u8");
    vslam::Frame next;
    next.status = vslam::Frame::Tracking;
    next.intensity_image_left = left1.data();
    next.intensity_image_right = right1.data();
    next.image_step = 1241;
    generator.initialize(&next);
    const double tx = -(386.1448 / 718.856) / 4;   // the band world moves the camera a quarter baseline per frame
    const std::array<double, 12> motion{{1, 0, 0, tx, 0, 1, 0, 0, 0, 0, 1, 0}};
    std::vector<int32_t> lost;
    generator.setProjectionTrackingDistancePixels(15);
    generator.setMaximumDescriptorDistanceTracking(25.6);
    generator.track(&next, &frame, motion, lost, false);
    generator.compute(&next);
    unsigned long long ht = 1469598103934665603ull;
    for (const vslam_track& t : next.tracks) {
      const int32_t v[5] = {t.index_previous, t.index_left, t.index_right, t.distance, t.epipolar_offset};
      for (int32_t x : v) ht = (ht ^ (unsigned long long)(uint32_t)x) * 1099511628211ull;
    }
    for (int32_t x : lost) ht = (ht ^ (unsigned long long)(uint32_t)x) * 1099511628211ull;
    for (const vslam_framepoint& p : next.points) ht = (ht ^ (unsigned long long)(uint32_t)p.index_left) * 1099511628211ull;
    std::printf("tracking %zu lost %zu landmarks %d new %zu average %.17g hash %llu\n", next.tracks.size(), lost.size(),
                generator.numberOfTrackedLandmarks(), next.points.size(), frame.average_descriptor_distance_tracking, ht);
    std::vector<vslam_previous_point> lost_points;
    for (int32_t i : lost) lost_points.push_back(frame.previous_points[i]);
    next.world_to_camera_left = motion;             // the first frame's camera is the world frame
    generator.setMaximumDescriptorDistanceTracking(64);
    generator.recoverPoints(&next, lost_points);
    unsigned long long hr = 1469598103934665603ull;
    for (const vslam_recovered_point& r : next.recovered) {
      hr = (hr ^ (unsigned long long)(uint32_t)r.index_lost) * 1099511628211ull;
      for (uint8_t b : r.descriptor_left) hr = (hr ^ b) * 1099511628211ull;
    }
    std::printf("recovered %zu hash %llu\n", next.recovered.size(), hr);

    // ---- StereoUVAligner::converge on the correspondence set written by the pytest
    const std::vector<uint8_t> raw = slurp(dir + "/correspondences.f64");
    const double* d = reinterpret_cast<const double*>(raw.data());
    const int32_t n = (int32_t)d[0];
    const double* moving = d + 1;
    const double* fixed = moving + 3 * (size_t)n;
    const double* omega = fixed + 4 * (size_t)n;
    const double* wt = omega + n;
    vslam_aligner_parameters ap = {1e-3, 16.0, 0.0, 1000, 0};   // configuration_kitti_fast.yaml:109-113
    vslam::StereoUVAligner aligner(ap, n);
    const double K[9] = {718.856, 0, 607.1928, 0, 718.856, 185.2157, 0, 0, 1};
    const double baseline[3] = {-386.1448, 0, 0};
    aligner.initialize(n, moving, fixed, omega, wt, K, baseline, 376, 1241, 0.1, {{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}});
    aligner.converge();
    std::printf("aligner converged %d rounds %d inliers %d outliers %d\n", (int)aligner.hasSystemConverged(),
                aligner.numberOfRounds(), aligner.numberOfInliers(), aligner.numberOfOutliers());
    const std::array<double, 12>& T = aligner.previousToCurrent();
    std::printf("pose");
    for (double v : T) std::printf(" %.17g", v);
    std::printf("\n");
    // ---- SURVEY 8f row 4 through the same header: the landmark loop of PoseTracker3D::_updatePoints as one call, and
    // the KITTI / TUM trajectory files of WorldMap (world_map.cpp:183-252)
    {
      const std::vector<uint8_t> lm = slurp(dir + "/landmarks.bin");
      const uint8_t* q = lm.data();
      int32_t n_landmarks, n_measurements, n_frames;
      std::memcpy(&n_landmarks, q, 4); std::memcpy(&n_measurements, q + 4, 4); std::memcpy(&n_frames, q + 8, 4);
      q += 16;
      std::vector<int32_t> offsets(n_landmarks + 1);
      std::memcpy(offsets.data(), q, 4 * offsets.size()); q += 4 * offsets.size() + (offsets.size() % 2 ? 4 : 0);
      std::vector<vslam_landmark_measurement> ms(n_measurements);
      std::memcpy(ms.data(), q, sizeof(vslam_landmark_measurement) * ms.size()); q += sizeof(vslam_landmark_measurement) * ms.size();
      std::vector<double> w2c(12 * (size_t)n_frames), c2w(12 * (size_t)n_frames), world(3 * (size_t)n_landmarks);
      std::memcpy(w2c.data(), q, 8 * w2c.size()); q += 8 * w2c.size();
      std::memcpy(c2w.data(), q, 8 * c2w.size()); q += 8 * c2w.size();
      std::memcpy(world.data(), q, 8 * world.size()); q += 8 * world.size();
      std::vector<uint32_t> updates(n_landmarks);
      std::memcpy(updates.data(), q, 4 * updates.size());
      vslam::LandmarkOptimizer optimizer(n_landmarks, n_measurements, n_frames);
      std::vector<uint8_t> outcome;
      optimizer.update(offsets, ms, w2c, c2w, world, updates, &outcome);
      unsigned long long hl = 1469598103934665603ull;
      for (double v : world) { unsigned long long bits; std::memcpy(&bits, &v, 8); hl = (hl ^ bits) * 1099511628211ull; }
      for (uint32_t u : updates) hl = (hl ^ u) * 1099511628211ull;
      for (uint8_t o : outcome) hl = (hl ^ o) * 1099511628211ull;
      std::printf("landmarks %d hash %llu\n", n_landmarks, hl);
      std::vector<double> timestamps(n_frames);
      for (int i = 0; i < n_frames; ++i) timestamps[i] = 1403636579.763555527 + 0.05 * i;
      vslam::writeTrajectoryKITTI(dir + "/trajectory_kitti.txt", c2w);
      vslam::writeTrajectoryTUM(dir + "/trajectory_tum.txt", timestamps, c2w);
      std::printf("trajectories %d\n", n_frames);
    }
    // error behaviour: exceptions, like the reference
    try {
      generator.initialize(nullptr);
      std::printf("no exception\n");
    } catch (const std::runtime_error& e) {
      std::printf("exception %s\n", e.what());
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "FAILED: %s\n", e.what());
    return 1;
  }
  return 0;
}
