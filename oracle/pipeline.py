"""Reference-owned driver logic of the hot path on top of tier A (C restatement) or tier B (cv2) primitives.

TEST INFRASTRUCTURE ONLY (see oracle/c/vslam_oracle.h).  Mirrors, call for call,
StereoFramePointGenerator::{configure, initialize, compute}
(/root/reference/src/framepoint_generation/stereo_framepoint_generator.cpp:16-60,73-133,135-462) and
BaseFramePointGenerator::{configure, detectKeypoints, computeDescriptors, adjustDetectorThresholds}
(/root/reference/src/framepoint_generation/base_framepoint_generator.cpp:165-329,355-459).
"""
from __future__ import annotations

import numpy as np

from . import tier_a


def _cv2():
    import cv2
    cv2.setNumThreads(0)  # executables/app.cpp:96
    return cv2


class StereoFramePointGeneratorOracle:
    def __init__(self, cfg, cam, tier: str = "a"):
        self.cfg, self.cam, self.tier = cfg, cam, tier
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
            a["x"] = [k.pt[0] + q["x"] for k in kps]
            a["y"] = [k.pt[1] + q["y"] for k in kps]
            a["response"] = [k.response for k in kps]
            out.append(a)
        return np.concatenate(out), np.asarray(counts, np.int32)

    def _describe(self, img, kps):
        """computeDescriptors (:431-438) -> (filtered kps, desc)"""
        if self.tier == "a":
            return tier_a.orb_compute(img, kps)
        cv2 = _cv2()
        cvk = [cv2.KeyPoint(float(k["x"]), float(k["y"]), 7.0, -1.0, float(k["response"]), 0, -1) for k in kps]
        cvk, desc = self._orb.compute(img, cvk)
        a = np.zeros(len(cvk), tier_a.KP)
        a["x"] = [k.pt[0] for k in cvk]
        a["y"] = [k.pt[1] for k in cvk]
        a["response"] = [k.response for k in cvk]
        return a, (desc if desc is not None else np.zeros((0, 32), np.uint8))

    # ---- stereo_framepoint_generator.cpp:73-133 ------------------------------------------------
    def initialize(self, left, right, localizing: bool):
        c = self.cfg
        self.thresholds_used = self.thresholds.copy()
        kl, cl = self._detect(left)
        kr, cr = self._detect(right)
        self.thresholds = tier_a.adjust_thresholds(                                          # :94
            self.thresholds, cl, cr, self.target_per_detector, c.target_number_of_keypoints_tolerance,
            c.detector_threshold_maximum_change, c.detector_threshold_minimum, c.detector_threshold_maximum)
        self.kps_left, self.desc_left = self._describe(left, kl)                             # :97-100
        self.kps_right, self.desc_right = self._describe(right, kr)
        self.counts_left, self.counts_right = cl, cr
        self.number_of_detected_keypoints = len(self.kps_left)                               # :101
        self.max_distance = tier_a.triangulation_threshold(                                  # :109-125
            localizing, self.number_of_detected_keypoints, self.target_number_of_keypoints,
            c.maximum_matching_distance_triangulation)
        self.features_left = tier_a.make_features(self.kps_left, self.desc_left)             # :129-132, :159-160
        self.features_right = tier_a.make_features(self.kps_right, self.desc_right)
        return self

    # ---- stereo_framepoint_generator.cpp:135-462 -----------------------------------------------
    def compute(self, tracked=None):
        c = self.cfg
        r = tier_a.stereo_compute(self.features_left, self.features_right, self.stereo_camera, self.max_distance,
                                  c.minimum_disparity_pixels, c.maximum_epipolar_search_offset_pixels,
                                  c.enable_keypoint_binning, c.bin_size_pixels, self.rows, self.cols, tracked)
        self.matches, self.winners = r["matches"], r["winners"]
        return r

    def framepoints(self):
        """new framepoints appended to frame->points() (only new matches; pre-loaded ones excluded)"""
        w = self.winners[self.winners >= 0]
        return self.matches[w]
