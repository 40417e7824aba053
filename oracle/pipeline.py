"""Reference-owned driver logic of the hot path on top of tier A (C restatement) or tier B (cv2) primitives.

TEST INFRASTRUCTURE ONLY (see oracle/c/vslam_oracle.h).  Mirrors, call for call,
StereoFramePointGenerator::{configure, initialize, compute}
(/root/reference/src/framepoint_generation/stereo_framepoint_generator.cpp:16-60,73-133,135-462) and
BaseFramePointGenerator::{configure, detectKeypoints, computeDescriptors, adjustDetectorThresholds}
(/root/reference/src/framepoint_generation/base_framepoint_generator.cpp:165-329,355-459).
"""
from __future__ import annotations

import time

import numpy as np

from . import tier_a


def _cv2():
    import cv2
    cv2.setNumThreads(0)  # executables/app.cpp:96
    return cv2


class StereoFramePointGeneratorOracle:
    def __init__(self, cfg, cam, tier: str = "a"):
        self.cfg, self.cam, self.tier = cfg, cam, tier
        # accumulated wall time per stage, named like the reference's chronometers (base_framepoint_generator.h:232-233,
        # stereo_framepoint_generator.h:81)
        self.seconds = {"keypoint_detection": 0.0, "descriptor_extraction": 0.0, "point_triangulation": 0.0}
        self.configure()

    # base_framepoint_generator.cpp:165-329 + stereo_framepoint_generator.cpp:16-60
    def configure(self):
        c, cam = self.cfg, self.cam
        self.rows, self.cols = cam.rows, cam.cols
        self.nv, self.nh = c.number_of_detectors_vertical, c.number_of_detectors_horizontal
        self.regions = tier_a.detector_regions(self.rows, self.cols, self.nv, self.nh)
        self.rows_bin, self.cols_bin = tier_a.bin_grid(self.rows, self.cols, c.bin_size_pixels)
        self.target_number_of_keypoints = self.rows_bin * self.cols_bin                      # :308
        self.target_per_detector = self.target_number_of_keypoints // (self.nv * self.nh)    # :312 (Count)
        self.thresholds = np.full(self.nv * self.nh, float(np.rint(c.detector_threshold_minimum)))  # :244, :13
        self.stereo_camera = tier_a.StereoCamera(cam.fx, cam.fy, cam.cx, cam.cy, cam.bx)
        if -cam.bx / cam.fx <= 0:                                                            # stereo :28-34
            raise RuntimeError("StereoFramePointGenerator::configure|invalid baseline")
        if self.tier == "b":
            cv2 = _cv2()
            self._orb = cv2.ORB_create()                                                     # :195,222

    # ---- primitives ---------------------------------------------------------------------------
    def _detect(self, img):
        """detectKeypoints (:355-429) -> (kps in reference order, raw counts per region)"""
        if self.tier == "a":
            return tier_a.detect_keypoints(img, self.nv, self.nh, self.thresholds)
        cv2 = _cv2()
        out, counts = [], []
        for q, thr in zip(self.regions, self.thresholds):
            det = cv2.FastFeatureDetector_create(int(np.rint(thr)))
            kps = det.detect(img[q["y"]:q["y"] + q["h"], q["x"]:q["x"] + q["w"]])
            counts.append(len(kps))
            a = np.zeros(len(kps), tier_a.KP)
            if len(kps):   # bulk conversions: the Python glue must not dominate the timed CPU baseline
                pts = cv2.KeyPoint_convert(kps)
                a["x"] = pts[:, 0] + np.float32(q["x"])                                      # :418-419
                a["y"] = pts[:, 1] + np.float32(q["y"])
                a["response"] = np.fromiter((k.response for k in kps), np.float32, len(kps))
            out.append(a)
        return np.concatenate(out), np.asarray(counts, np.int32)

    def _describe(self, img, kps):
        """computeDescriptors (:431-438) -> (filtered kps, desc)"""
        if getattr(self.cfg, "descriptor_type", "ORB") == "BRIEF":   # :186 (no cv2 counterpart here: tier A only)
            return tier_a.brief32_compute(img, kps, self.cfg.brief_tests)
        if self.tier == "a":
            return tier_a.orb_compute(img, kps)
        cv2 = _cv2()
        if len(kps) == 0:
            return kps, np.zeros((0, 32), np.uint8)
        pts = np.ascontiguousarray(np.stack([kps["x"], kps["y"]], 1))
        cvk = cv2.KeyPoint_convert(pts, size=7.0)            # angle -1, octave 0, class_id -1 like FAST's own
        cvk, desc = self._orb.compute(img, cvk)
        if desc is None:
            return kps[:0], np.zeros((0, 32), np.uint8)
        kept = cv2.KeyPoint_convert(cvk)
        # ORB only drops keypoints (31 px border) and keeps the order: carry the responses over by position
        h, w = img.shape
        keep = (kps["x"] >= 31) & (kps["x"] < w - 31) & (kps["y"] >= 31) & (kps["y"] < h - 31)
        a = kps[keep].copy()
        assert len(a) == len(kept) and np.array_equal(a["x"], kept[:, 0]) and np.array_equal(a["y"], kept[:, 1])
        return a, desc

    # ---- stereo_framepoint_generator.cpp:73-133 ------------------------------------------------
    def initialize(self, left, right, localizing: bool):
        c = self.cfg
        self.thresholds_used = self.thresholds.copy()
        t0 = time.perf_counter()             # CHRONOMETER keypoint_detection (base_framepoint_generator.h:232-233)
        kl, cl = self._detect(left)
        kr, cr = self._detect(right)
        self.thresholds = tier_a.adjust_thresholds(                                          # :94
            self.thresholds, cl, cr, self.target_per_detector, c.target_number_of_keypoints_tolerance,
            c.detector_threshold_maximum_change, c.detector_threshold_minimum, c.detector_threshold_maximum)
        t1 = time.perf_counter()             # CHRONOMETER descriptor_extraction
        self.kps_left, self.desc_left = self._describe(left, kl)                             # :97-100
        self.kps_right, self.desc_right = self._describe(right, kr)
        t2 = time.perf_counter()
        self.seconds["keypoint_detection"] += t1 - t0
        self.seconds["descriptor_extraction"] += t2 - t1
        self.counts_left, self.counts_right = cl, cr
        self.number_of_detected_keypoints = len(self.kps_left)                               # :101
        self.max_distance = tier_a.triangulation_threshold(                                  # :109-125
            localizing, self.number_of_detected_keypoints, self.target_number_of_keypoints,
            c.maximum_matching_distance_triangulation)
        self.features_left = tier_a.make_features(self.kps_left, self.desc_left)             # :129-132, :159-160
        self.features_right = tier_a.make_features(self.kps_right, self.desc_right)
        self.images = (left, right)
        self._blurred = None
        return self

    # ---- stereo_framepoint_generator.cpp:464-681 -----------------------------------------------
    def track(self, previous, previous_to_current, track_by_appearance, tracking_distance_pixels,
              max_distance_tracking):
        """previous: tier_a.PREVIOUS_POINT records of frame_previous->points().  Prunes the matched features from the
        candidate pools like :671-672, so that the next compute() only scans the rest."""
        c = self.cfg
        r = tier_a.track(self.features_left, self.features_right, self.rows, self.cols, self.stereo_camera, previous,
                         previous_to_current, track_by_appearance, tracking_distance_pixels, max_distance_tracking,
                         self.max_distance, c.minimum_disparity_pixels)
        self.features_left = self.features_left[~r["matched_left"]]
        self.features_right = self.features_right[~r["matched_right"]]
        self.tracks = r["tracks"]
        return r

    @staticmethod
    def tracked_points(tracks):
        """the tracked points as compute() finds them in frame->points() (:147-155): FramePoint row/col
        (frame_point.cpp:8-24), previous() set, disparityPixels, descriptorDistanceTriangulation"""
        t = np.zeros(len(tracks), tier_a.TRACKED)
        t["row"], t["col"] = tracks["yl"].astype(np.int32), tracks["xl"].astype(np.int32)
        t["has_previous"] = 1
        t["disparity"] = (tracks["xl"] - tracks["xr"]).astype(np.float64)
        t["distance"] = tracks["distance"]
        return t

    # ---- stereo_framepoint_generator.cpp:683-869 -----------------------------------------------
    def recover_points(self, lost, world_to_camera_left, max_distance_tracking, min_depth=0.1, max_depth=1000.0):
        """lost: tier_a.PREVIOUS_POINT records of the lost points (parameters.h:196-199 depth defaults)"""
        if getattr(self.cfg, "descriptor_type", "ORB") == "BRIEF":
            return tier_a.recover_points(self.images[0], self.images[1], self.stereo_camera, lost, world_to_camera_left,
                                         min_depth, max_depth, max_distance_tracking, self.max_distance,
                                         self.cfg.minimum_disparity_pixels, self.cfg.brief_tests)
        if self._blurred is None:
            self._blurred = (tier_a.gauss7_u8(self.images[0]), tier_a.gauss7_u8(self.images[1]))
        return tier_a.recover_points(self._blurred[0], self._blurred[1], self.stereo_camera, lost,
                                     world_to_camera_left, min_depth, max_depth, max_distance_tracking,
                                     self.max_distance, self.cfg.minimum_disparity_pixels)

    # ---- stereo_framepoint_generator.cpp:135-462 -----------------------------------------------
    def compute(self, tracked=None):
        c = self.cfg
        t0 = time.perf_counter()             # CHRONOMETER point_triangulation (stereo_framepoint_generator.h:81)
        r = tier_a.stereo_compute(self.features_left, self.features_right, self.stereo_camera, self.max_distance,
                                  c.minimum_disparity_pixels, c.maximum_epipolar_search_offset_pixels,
                                  c.enable_keypoint_binning, c.bin_size_pixels, self.rows, self.cols, tracked)
        self.seconds["point_triangulation"] += time.perf_counter() - t0
        self.matches, self.winners = r["matches"], r["winners"]
        return r

    # ---- stereo_framepoint_generator.cpp:168-273 ------------------------------------------------
    def dead_use_matches_block(self):
        """The block `use_matches: true` (the struct default, parameters.h:224-237; kitti_fast / euroc YAMLs do not
        override it) executes at the top of compute(): CV_32F conversion of both descriptor matrices, FLANN knnMatch
        (k = 2) and cv::findHomography(RANSAC, reproject 1, 1000 iterations, confidence 0.99).  Its results
        (left_good_points / right_good_points) are never read: it only costs time.  Tier B (cv2) only; used by
        bench.py to time the reference's CPU path *as configured*."""
        cv2 = _cv2()
        if len(self.desc_left) < 4 or len(self.desc_right) < 4:
            return 0
        d1, d2 = self.desc_left.astype(np.float32), self.desc_right.astype(np.float32)          # :199-205
        matcher = cv2.DescriptorMatcher_create(cv2.DescriptorMatcher_FLANNBASED)                 # :175-177
        knn = matcher.knnMatch(d1, d2, 2)                                                        # :206
        q = np.fromiter((m[0].queryIdx for m in knn), np.int64, len(knn))                        # :224-238
        t = np.fromiter((m[0].trainIdx for m in knn), np.int64, len(knn))
        p1 = np.stack([self.kps_left["x"][q], self.kps_left["y"][q]], 1).astype(np.float64)
        p2 = np.stack([self.kps_right["x"][t], self.kps_right["y"][t]], 1).astype(np.float64)
        _, mask = cv2.findHomography(p1, p2, cv2.RANSAC, 1.0, maxIters=1000, confidence=0.99)    # :242-246
        return 0 if mask is None else int(mask.sum())

    def framepoints(self):
        """new framepoints appended to frame->points() (only new matches; pre-loaded ones excluded)"""
        w = self.winners[self.winners >= 0]
        return self.matches[w]
