"""Tier A (C restatement) against tier B (cv2 4.13 primitives) on full-size synthetic inputs.
cv2 is the live cross-check here; the same boundary is also pinned offline by tests/golden/."""
import numpy as np
import pytest

from oracle import pipeline, tier_a
from vslam_b200 import configs, synth

cv2 = pytest.importorskip("cv2")


def test_gaussian_kernel_matches_cv2():
    assert np.array_equal(tier_a.gauss7_kernel(), cv2.getGaussianKernel(7, 2, cv2.CV_32F).ravel())


@pytest.mark.parametrize("shape,threshold", [("kitti", 20), ("kitti", 35), ("euroc", 10)])
def test_fast_matches_cv2(shape, threshold):
    left, _ = synth.band_world_pair(shape, 3)
    kps = cv2.FastFeatureDetector_create(threshold).detect(left)
    want = np.array([[k.pt[0], k.pt[1], k.response] for k in kps], np.float32)
    kp = tier_a.fast_detect(left, threshold)
    assert np.array_equal(np.stack([kp["x"], kp["y"], kp["response"]], 1), want)
    assert all(k.size == 7 and k.angle == -1 and k.octave == 0 and k.class_id == -1 for k in kps[:50])


def test_fast_on_roi_view_matches_cv2():
    left, _ = synth.band_world_pair("euroc", 5)
    view = left[238:480, 374:752]          # the (1,1) detector region of the 2x2 EuRoC grid
    kps = cv2.FastFeatureDetector_create(15).detect(view)
    kp = tier_a.fast_detect(view, 15)
    assert np.array_equal(np.stack([kp["x"], kp["y"]], 1), np.array([k.pt for k in kps], np.float32))


def test_blur_matches_float_separable_path():
    left, _ = synth.band_world_pair("kitti", 7)
    k = cv2.getGaussianKernel(7, 2, cv2.CV_32F).ravel()
    ref = np.rint(cv2.sepFilter2D(left, cv2.CV_32F, k, k, borderType=cv2.BORDER_REFLECT_101)).astype(np.uint8)
    assert np.array_equal(tier_a.gauss7_u8(left), ref)


@pytest.mark.parametrize("cfgname,seed,localizing", [("kitti", 1, False), ("kitti_fast", 4, True),
                                                     ("euroc", 0, True), ("euroc", 1, False)])
def test_full_framepoint_generation_tier_a_equals_tier_b(cfgname, seed, localizing):
    cfg = configs.BY_NAME[cfgname]
    cam = synth.camera(cfg.camera)
    left, right = synth.band_world_pair(cfg.camera, seed)
    a = pipeline.StereoFramePointGeneratorOracle(cfg, cam, "a").initialize(left, right, localizing)
    b = pipeline.StereoFramePointGeneratorOracle(cfg, cam, "b").initialize(left, right, localizing)
    a.compute()
    b.compute()
    assert np.array_equal(a.counts_left, b.counts_left) and np.array_equal(a.counts_right, b.counts_right)
    assert np.array_equal(a.thresholds, b.thresholds)
    assert np.array_equal(a.kps_left, b.kps_left) and np.array_equal(a.kps_right, b.kps_right)
    assert np.array_equal(a.desc_left, b.desc_left) and np.array_equal(a.desc_right, b.desc_right)
    assert np.array_equal(a.matches, b.matches) and np.array_equal(a.winners, b.winners)
    assert len(a.matches) > 500 and len(a.winners) > 300


def test_hamming_matches_cv2_norm():
    rng = np.random.default_rng(0)
    d = rng.integers(0, 256, (40, 32), dtype=np.uint8)
    for i in range(39):
        assert tier_a.hamming256(d[i], d[i + 1]) == int(cv2.norm(d[i], d[i + 1], cv2.NORM_HAMMING))
