"""Parity of the CUDA framepoint generator (through the C ABI) with the CPU oracle: bit-exact keypoint sets,
scores, order, descriptors, thresholds, match lists, bin winners; triangulated points bit-exact too (the kernel
uses the oracle's evaluation order with no contraction), asserted at 0 tolerance and documented as <= 1e-4 rel."""
import os

import numpy as np
import pytest

from oracle import pipeline, tier_a
from vslam_b200 import api, configs, synth

pytestmark = pytest.mark.gpu


def _oracle(cfg, cam, left, right, localizing, thresholds=None, tracked=None):
    o = pipeline.StereoFramePointGeneratorOracle(cfg, cam, "a")
    if thresholds is not None:
        o.thresholds = np.asarray(thresholds, np.float64).copy()
    o.initialize(left, right, localizing)
    o.compute(None if tracked is None else tracked.astype(tier_a.TRACKED))
    return o


def _same_points(got, want):
    assert len(got) == len(want)
    for f, g in (("index_left", "index_left"), ("index_right", "index_right"), ("xl", "xl"), ("yl", "yl"),
                 ("xr", "xr"), ("yr", "yr"), ("distance", "distance"), ("epipolar_offset", "epipolar_offset")):
        assert np.array_equal(got[f], want[g]), f
    # tolerance stated by north_star: 1e-4 relative; achieved: bit-exact
    assert np.array_equal(got["camera"], want["cam"])


def _check_pair(gen, o, side_features=True):
    for side, (kps, desc) in enumerate(((o.kps_left, o.desc_left), (o.kps_right, o.desc_right))):
        k, d = gen.features(side)
        assert len(k) == len(kps)
        assert np.array_equal(k["x"], kps["x"]) and np.array_equal(k["y"], kps["y"])
        assert np.array_equal(k["response"], kps["response"])
        assert np.array_equal(d, desc)


@pytest.mark.parametrize("cfgname,seed,localizing", [("kitti", 1, True), ("kitti", 2, False), ("kitti_fast", 4, True),
                                                     ("euroc", 0, True), ("euroc", 1, False), ("hd", 0, True)])
def test_initialize_and_compute_match_oracle(cfgname, seed, localizing):
    cfg = configs.BY_NAME[cfgname]
    cam = synth.camera(cfg.camera)
    left, right = synth.band_world_pair(cfg.camera, seed)
    gen = api.StereoFramePointGenerator(cfg, cam)
    nl, nr = gen.initialize(left, right, localizing)
    o = _oracle(cfg, cam, left, right, localizing)
    assert (nl, nr) == (len(o.kps_left), len(o.kps_right))
    cl, cr, dist = gen.detection_stats()
    assert np.array_equal(cl, o.counts_left) and np.array_equal(cr, o.counts_right)
    assert dist == o.max_distance
    assert np.array_equal(gen.thresholds, o.thresholds)
    _check_pair(gen, o)
    fps = gen.compute()
    assert gen.number_of_matches == len(o.matches)
    _same_points(fps, o.framepoints())
    _same_points(gen.matches(), o.matches)
    assert len(fps) > 300
    gen.close()


@pytest.mark.parametrize("cfgname", ["kitti", "euroc"])
def test_raw_fast_mask_and_blur_taps_match_oracle(cfgname):
    cfg = configs.BY_NAME[cfgname]
    cam = synth.camera(cfg.camera)
    left, right = synth.band_world_pair(cfg.camera, 9)
    gen = api.StereoFramePointGenerator(cfg, cam)
    thr = gen.thresholds
    gen.initialize(left, right, True)
    for side, img in enumerate((left, right)):
        kps, _ = tier_a.detect_keypoints(img, cfg.number_of_detectors_vertical, cfg.number_of_detectors_horizontal, thr)
        want = np.zeros((cam.rows, cam.cols), bool)
        want[kps["y"].astype(int), kps["x"].astype(int)] = True
        assert np.array_equal(gen.debug_keypoint_mask(side), want)      # every raw FAST keypoint, border included
        assert np.array_equal(gen.debug_blurred(side), tier_a.gauss7_u8(img))
    gen.close()


def test_sequence_threshold_controller_feedback_matches_oracle():
    """frames of one sequence: the thresholds of frame t+1 depend on the counts of frame t (L and R)."""
    cfg, cam = configs.KITTI_FAST, synth.camera("kitti")
    world = synth.BandWorld(cam.cols, cam.rows, 1, max_frames=8)
    gen = api.StereoFramePointGenerator(cfg, cam)
    o = pipeline.StereoFramePointGeneratorOracle(cfg, cam, "a")
    seen = []
    for k in range(6):
        left, right = world.pair(k)
        gen.initialize(left, right, k == 0)
        o.initialize(left, right, k == 0)
        o.compute()
        assert np.array_equal(gen.thresholds, o.thresholds), k
        _check_pair(gen, o)
        _same_points(gen.compute(), o.framepoints())
        seen.append(float(o.thresholds[0]))
    assert len(set(seen)) > 2          # the controller really moved
    gen.close()


@pytest.mark.parametrize("exact_in_float", [True, False])
def test_tracked_points_preload_bins(exact_in_float):
    """exact_in_float=False sends disparities that do not round-trip through float: the generic select kernel"""
    cfg, cam = configs.KITTI, synth.camera("kitti")
    left, right = synth.band_world_pair("kitti", 3)
    gen = api.StereoFramePointGenerator(cfg, cam)
    gen.initialize(left, right, False)
    base = gen.compute()
    # pretend every third winner had been tracked from the previous frame (has_previous), plus two untracked
    tracked = np.zeros(len(base[::3]) + 2, api.TRACKED)
    tracked["row"][:-2] = base["yl"][::3].astype(int)
    tracked["col"][:-2] = base["xl"][::3].astype(int)
    tracked["has_previous"][:-2] = 1
    tracked["row"][-2:], tracked["col"][-2:] = (100, 200), (500, 900)
    tracked["disparity"][-2:], tracked["distance"][-2:] = (1000.0, 0.5), (0.0, 300.0)
    if not exact_in_float:
        tracked["disparity"][:-2] = 0.1
    o = _oracle(cfg, cam, left, right, False, tracked=tracked)
    gen.close()
    gen = api.StereoFramePointGenerator(cfg, cam)     # compute() consumes the matched features, like the reference
    gen.initialize(left, right, False)
    fps = gen.compute(tracked)
    w = o.winners
    assert len(fps) == len(w) and (w < 0).sum() >= 1
    new = w >= 0
    _same_points(fps[new], o.matches[w[new]])
    assert np.array_equal(fps["index_left"][~new], w[~new])       # surviving pre-loaded point k reported as -(k+1)
    assert len(fps) < len(base)
    gen.close()


def test_epipolar_offset_three_like_test_stereo_frontend():
    """executables/test_stereo_frontend.cpp:98-104: pinned threshold, offset 3 (7 passes), distance 35."""
    import dataclasses
    cfg = dataclasses.replace(configs.KITTI, detector_threshold_minimum=25, detector_threshold_maximum=25,
                              maximum_matching_distance_triangulation=35.0, maximum_epipolar_search_offset_pixels=3)
    cam = synth.camera("kitti")
    left, right = synth.band_world_pair("kitti", 5)
    right = np.roll(right, 1, axis=0).copy()      # true matches now sit one row lower in the right image
    right[:2] = 96
    gen = api.StereoFramePointGenerator(cfg, cam)
    gen.initialize(left, right, False)
    o = _oracle(cfg, cam, left, right, False)
    fps = gen.compute()
    assert gen.number_of_matches == len(o.matches)      # matches of EVERY epipolar pass are counted (:397-398)
    _same_points(gen.matches(), o.matches)
    _same_points(fps, o.framepoints())
    offs = set(o.matches["epipolar_offset"].tolist())
    assert -1 in offs and len(offs) >= 3
    gen.close()


def test_binning_disabled_returns_every_match_in_emission_order():
    import dataclasses
    cfg = dataclasses.replace(configs.EUROC, enable_keypoint_binning=False)
    cam = synth.camera("euroc")
    left, right = synth.band_world_pair("euroc", 2)
    gen = api.StereoFramePointGenerator(cfg, cam)
    gen.initialize(left, right, True)
    o = _oracle(cfg, cam, left, right, True)
    fps = gen.compute()
    assert len(fps) == len(o.matches) == gen.number_of_matches
    _same_points(fps, o.matches)
    gen.close()


@pytest.mark.parametrize("cfgname,n", [("kitti_fast", 7), ("euroc", 5), ("hd", 3)])
def test_batched_pairs_equal_independent_first_frames(cfgname, n):
    cfg = configs.BY_NAME[cfgname]
    cam = synth.camera(cfg.camera)
    left, right = synth.band_world_batch(cfg.camera, range(20, 20 + n))
    os.environ["VSLAM_CHUNK_PAIRS"] = "3"          # force several chunks on both pipeline lanes
    try:
        gen = api.StereoFramePointGenerator(cfg, cam, max_batch=n)
    finally:
        del os.environ["VSLAM_CHUNK_PAIRS"]
    out, counts = gen.batch_process(left, right, True)
    gen.batch_upload(left, right)
    gen.batch_run(n, True)
    out2, nf, nm, nl, nr = gen.batch_download(n)
    assert np.array_equal(counts, nf)
    for i in range(n):
        o = _oracle(cfg, cam, left[i], right[i], True)
        _same_points(out[i, :counts[i]], o.framepoints())
        _same_points(out2[i, :nf[i]], o.framepoints())
        assert nm[i] == len(o.matches) and nl[i] == len(o.kps_left) and nr[i] == len(o.kps_right)
        k, d = gen.features(0, pair=i)
        assert np.array_equal(d, o.desc_left) and np.array_equal(k["x"], o.kps_left["x"])
    assert np.array_equal(gen.thresholds, np.full(gen.number_of_detectors, cfg.detector_threshold_minimum))
    gen.close()


def test_batched_pairs_with_epipolar_offsets_count_every_pass():
    """maximum_epipolar_search_offset_pixels > 0 in the batched (strip-select) path: number_of_new_points sums the
    matches of every pass (stereo_framepoint_generator.cpp:278, :397-398)."""
    import dataclasses
    cfg = dataclasses.replace(configs.KITTI_FAST, maximum_epipolar_search_offset_pixels=2)
    cam = synth.camera(cfg.camera)
    n = 3
    left, right = synth.band_world_batch(cfg.camera, range(60, 60 + n))
    right = np.roll(right, 1, axis=1).copy()       # true matches sit one row lower in the right images
    right[:, :2] = 96
    # (a keypoint capacity that is not a multiple of 4: the per-pair shares of the byte-flag arrays are then unaligned)
    gen = api.StereoFramePointGenerator(cfg, cam, max_batch=n, max_keypoints=4099)
    out, counts = gen.batch_process(left, right, True)
    _, nf, nm, _, _ = gen.batch_download(n)
    later = 0
    for i in range(n):
        o = _oracle(cfg, cam, left[i], right[i], True)
        assert nm[i] == len(o.matches) and nf[i] == counts[i]
        _same_points(out[i, :counts[i]], o.framepoints())
        later += int((o.matches["epipolar_offset"] != 0).sum())
    assert later > 100
    gen.close()


@pytest.mark.parametrize("name", ["kitti_crop", "euroc_crop"])
def test_against_committed_cv2_golden_vectors(golden_dir, name):
    import dataclasses
    g = np.load(os.path.join(golden_dir, "fast_orb_%s.npz" % name))
    img = g["image"]
    cam = synth.Camera(img.shape[1], img.shape[0], 400.0, 400.0, img.shape[1] / 2, img.shape[0] / 2, -40.0)
    cfg = dataclasses.replace(configs.KITTI, detector_threshold_minimum=12, detector_threshold_maximum=12)
    gen = api.StereoFramePointGenerator(cfg, cam)
    gen.initialize(img, img, True)
    want = np.zeros(img.shape, bool)
    want[g["fast12"][:, 1].astype(int), g["fast12"][:, 0].astype(int)] = True
    assert np.array_equal(gen.debug_keypoint_mask(0), want)
    k, d = gen.features(0)
    assert np.array_equal(np.stack([k["x"], k["y"], k["response"]], 1), g["orb_kps"])
    assert np.array_equal(d, g["orb_desc"])
    gen.close()


def test_edge_cases_blank_tiny_and_capacity():
    import dataclasses
    cam = synth.Camera(96, 80, 100.0, 100.0, 48.0, 40.0, -10.0)
    gen = api.StereoFramePointGenerator(configs.KITTI, cam)
    blank = np.full((80, 96), 128, np.uint8)
    assert gen.initialize(blank, blank, True) == (0, 0)
    assert len(gen.compute()) == 0 and gen.number_of_matches == 0
    assert gen.thresholds[0] == 20.0            # already at the minimum
    gen.close()
    # an image smaller than the 31 px descriptor border on each side: raw corners but no descriptors
    cam = synth.Camera(40, 40, 100.0, 100.0, 20.0, 20.0, -10.0)
    rng = np.random.default_rng(0)
    noise = rng.integers(0, 255, (40, 40), dtype=np.uint8)
    gen = api.StereoFramePointGenerator(configs.KITTI, cam)
    assert gen.initialize(noise, noise, True) == (0, 0)
    cl, _, _ = gen.detection_stats()
    kps, counts = tier_a.detect_keypoints(noise, 1, 1, [20.0])
    assert cl[0] == counts[0] > 0
    gen.close()
    # capacity overflow is an error, not a truncation
    kcam = synth.camera("kitti")
    left, right = synth.band_world_pair("kitti", 1)
    gen = api.StereoFramePointGenerator(configs.KITTI, kcam, max_keypoints=500)
    with pytest.raises(api.VslamError) as e:
        gen.initialize(left, right, True)
    assert e.value.code == -3
    gen.close()


@pytest.mark.parametrize("extra_columns", [40, 1000])
def test_row_strided_input_views(extra_columns):
    """images that are views into wider buffers: a moderate row stride is uploaded by one linear copy + repitch (the
    padding travels too), a wide one by a strided 2-D copy"""
    cfg, cam = configs.EUROC, synth.camera("euroc")
    left, right = synth.band_world_pair("euroc", 6)
    big_l = np.full((cam.rows, cam.cols + extra_columns), 77, np.uint8)
    big_r = np.full_like(big_l, 201)
    big_l[:, :cam.cols], big_r[:, :cam.cols] = left, right
    gen = api.StereoFramePointGenerator(cfg, cam)
    gen.initialize(big_l[:, :cam.cols], big_r[:, :cam.cols], True)
    o = _oracle(cfg, cam, left, right, True)
    _check_pair(gen, o)
    gen.close()


def test_batched_self_alignment_linearize_matches_oracle():
    """vslam_fpg_batch_linearize: per pair, StereoUVAligner::initialize + linearize over the pair's own framepoints."""
    cfg, acfg = configs.KITTI_FAST, configs.KITTI_FAST_ALIGNER
    cam = synth.camera(cfg.camera)
    n = 4
    left, right = synth.band_world_batch(cfg.camera, range(40, 40 + n))
    gen = api.StereoFramePointGenerator(cfg, cam, max_batch=n)
    out, counts = gen.batch_process(left, right, True)
    T = synth.true_motion(0.15)             # a small prior error so that inliers and outliers both occur
    seen_in = seen_out = 0
    for ignore in (False, True):
        gen.batch_linearize(n, T, acfg, ignore_outliers=ignore, rounds=2)
        systems, errors, inliers = gen.batch_systems(n, with_points=True)
        for i in range(n):
            fp = out[i, :counts[i]]
            moving = fp["camera"]
            fixed = np.stack([fp["xl"], fp["yl"], fp["xr"], fp["yr"]], 1).astype(np.float64)
            wt = np.minimum(acfg.maximum_reliable_depth_meters / moving[:, 2], 1.0)
            ora = tier_a.Aligner("stereouv", moving, fixed, np.ones(len(fp)), wt, cam.K, cam.baseline, cam.rows, cam.cols,
                                 acfg.minimum_reliable_depth_meters, acfg.maximum_error_kernel)
            want = ora.linearize(T, ignore)
            got = systems[i]
            assert got["inliers"] == want["inliers"] and got["outliers"] == want["outliers"]
            seen_in += want["inliers"]
            seen_out += want["outliers"]
            np.testing.assert_allclose(got["H"], want["H"], rtol=1e-10, atol=1e-12 * np.abs(want["H"]).max())
            np.testing.assert_allclose(got["b"], want["b"], rtol=1e-9, atol=1e-12 * np.abs(want["H"]).max())
            np.testing.assert_allclose(got["total_error"], want["total_error"], rtol=1e-12)
            assert np.array_equal(errors[i, :counts[i]], ora.errors)
            assert np.array_equal(inliers[i, :counts[i]], ora.inliers)
    assert seen_in > 100 and seen_out > 100
    gen.close()


def test_single_pair_initialize_runs_as_a_cuda_graph():
    """the device side of initialize() is one CUDA-graph launch per frame (captured once, re-captured when the row
    stride or the profiling mode changes; with profiling the timing events are external event-record nodes of the
    graph) -- the results are those of the kernels launched one by one (every other test's oracle comparison)"""
    cfg, cam = configs.KITTI, synth.camera("kitti")
    gen = api.StereoFramePointGenerator(cfg, cam)
    ref = api.StereoFramePointGenerator(cfg, cam)
    ref.set_profiling(True)
    for k in range(3):
        left, right = synth.band_world_pair("kitti", 30 + k)
        assert gen.initialize(left, right, k == 0) == ref.initialize(left, right, k == 0)
        for side in (0, 1):
            (ka, da), (kb, db) = gen.features(side), ref.features(side)
            assert np.array_equal(ka, kb) and np.array_equal(da, db)
        assert np.array_equal(gen.compute(), ref.compute())
        assert np.array_equal(gen.thresholds, ref.thresholds)
    assert gen.graph_launch_count == 3 and ref.graph_launch_count == 3
    assert all(v > 0 for v in ref.time_consumption().values())
    wide_l = np.zeros((cam.rows, cam.cols + 24), np.uint8)                  # another row stride: captured again
    wide_r = np.zeros_like(wide_l)
    wide_l[:, :cam.cols], wide_r[:, :cam.cols] = left, right
    assert gen.initialize(wide_l[:, :cam.cols], wide_r[:, :cam.cols], False) == ref.initialize(left, right, False)
    assert gen.graph_launch_count == 4 and ref.graph_launch_count == 4
    assert np.array_equal(gen.features(0)[1], ref.features(0)[1])
    gen.close()
    ref.close()


def test_chronometers_and_kernel_profile():
    cfg, cam = configs.KITTI, synth.camera("kitti")
    left, right = synth.band_world_pair("kitti", 1)
    gen = api.StereoFramePointGenerator(cfg, cam)
    gen.set_profiling(True)
    gen.initialize(left, right, True)
    gen.compute()
    t = gen.time_consumption()
    assert t["keypoint_detection"] > 0 and t["descriptor_extraction"] > 0 and t["point_triangulation"] > 0
    prof = gen.kernel_profile()
    assert all(prof[k][1] == 1 for k in ("fast_nms", "compact", "blur", "describe", "match", "select"))
    assert gen.launch_count == 8      # + repitch, + the pack kernel of the feature prefetch that initialize() starts
    gen.close()


def test_features_consumed_by_tracking_are_excluded_from_the_scan():
    """track() prunes the features it matched before compute() runs (stereo_framepoint_generator.cpp:671-672)."""
    cfg, cam = configs.KITTI, synth.camera("kitti")
    left, right = synth.band_world_pair("kitti", 8)
    gen = api.StereoFramePointGenerator(cfg, cam)
    gen.initialize(left, right, False)
    o = pipeline.StereoFramePointGeneratorOracle(cfg, cam, "a").initialize(left, right, False)
    rng = np.random.default_rng(0)
    keep_l = rng.random(len(o.kps_left)) > 0.3
    keep_r = rng.random(len(o.kps_right)) > 0.3
    gen.set_remaining_features(0, o.kps_left[keep_l])
    gen.set_remaining_features(1, o.kps_right[keep_r])
    # oracle: the matcher's feature vectors without the pruned features (indices keep referring to the frame's arrays)
    fl = o.features_left[keep_l[o.features_left["index"]]]
    fr = o.features_right[keep_r[o.features_right["index"]]]
    r = tier_a.stereo_compute(fl, fr, o.stereo_camera, o.max_distance, cfg.minimum_disparity_pixels, 0, True,
                              cfg.bin_size_pixels, cam.rows, cam.cols)
    fps = gen.compute()
    _same_points(gen.matches(), r["matches"])
    _same_points(fps, r["matches"][r["winners"]])
    assert 200 < len(r["matches"]) < len(o.kps_left) * 0.6
    with pytest.raises(api.VslamError):
        bad = np.zeros(1, api.KEYPOINT)
        bad["x"], bad["y"] = 5, 5
        gen.set_remaining_features(0, bad)
    gen.close()


def test_rows_with_more_than_32_features_take_the_general_scan_path():
    """dense texture: many image rows hold > 32 features, which leaves the register-resident fast path of match_kernel"""
    import dataclasses
    cam = synth.camera("kitti")
    rng = np.random.default_rng(5)
    canvas = rng.integers(25, 36, (cam.rows, cam.cols + 64)).astype(np.uint8)
    for y in range(36, cam.rows - 36, 9):        # isolated bright dots every ~7 px: each one is a FAST keypoint
        xs = np.arange(20, cam.cols + 44, 7)
        xs = xs + rng.integers(0, 2, len(xs))
        canvas[y, xs] = rng.integers(150, 251, len(xs))
    left = np.ascontiguousarray(canvas[:, 8:8 + cam.cols])
    right = np.ascontiguousarray(canvas[:, 14:14 + cam.cols])                  # 6 px disparity, ambiguous matches
    cfg = dataclasses.replace(configs.KITTI, detector_threshold_minimum=30, detector_threshold_maximum=30)
    gen = api.StereoFramePointGenerator(cfg, cam, max_keypoints=60000)
    gen.initialize(left, right, True)
    o = _oracle(cfg, cam, left, right, True)
    per_row = np.bincount(o.kps_left["y"].astype(int))
    assert per_row.max() > 100 and len(o.matches) > 5000
    _check_pair(gen, o)
    fps = gen.compute()
    _same_points(gen.matches(), o.matches)
    _same_points(fps, o.framepoints())
    gen.close()


@pytest.mark.parametrize("threshold", [0, 1, 2, 5, 127, 128, 200])
def test_fast_threshold_extremes(threshold):
    """t = 0 takes the exactly packed arc test, t >= 1 the multiply-add packing with the biased dark lane, t > 127 the
    path without the byte-SIMD pre-test: all must give cv::FAST's keypoints and scores"""
    cfg = configs.KITTI
    cam = synth.camera(cfg.camera)
    rng = np.random.default_rng(threshold)
    left, right = synth.band_world_pair(cfg.camera, 3)
    # low-contrast noise patches (differences of 0, 1, 2 around a flat level) and saturated blocks
    left = left.copy()
    left[40:140, 100:400] = 100 + rng.integers(0, 3, (100, 300))
    left[200:300, 500:900] = np.where(rng.random((100, 400)) < 0.5, 0, 255)
    gen = api.StereoFramePointGenerator(cfg, cam, max_keypoints=65535)
    gen.thresholds = [threshold]
    try:
        gen.initialize(left, right, True)
    except api.VslamError as e:          # t = 0 on noise: more keypoints than any handle can hold is a capacity error
        assert e.code == -3 and threshold == 0
        gen.close()
        return
    kps = tier_a.fast_detect(left, threshold)
    want = np.zeros((cam.rows, cam.cols), bool)
    want[kps["y"].astype(int), kps["x"].astype(int)] = True
    assert np.array_equal(gen.debug_keypoint_mask(0), want)
    k, _ = gen.features(0)
    keep = (kps["x"] >= 31) & (kps["x"] < cam.cols - 31) & (kps["y"] >= 31) & (kps["y"] < cam.rows - 31)
    ref = kps[keep]
    order = np.lexsort((ref["x"], ref["y"]))
    assert np.array_equal(k["x"], ref["x"][order]) and np.array_equal(k["response"], ref["response"][order])
    gen.close()


def test_full_size_batch_is_replication_invariant():
    """BASELINE configs[2] at its full size (4096 independent KITTI-shape pairs, kitti_fast): the oracle cannot run
    4096 pairs in seconds, so the size-independent property is checked instead -- a pair's result does not depend on
    its position in the batch, on the chunk or on the pipeline lane that processed it.  The batch tiles 8 distinct
    pairs 512 times; every replica must equal, byte for byte, the first occurrence, whose 8 results are compared with
    the oracle; the descriptors of the last replica (last chunk) are compared as well."""
    cfg = configs.BY_NAME["kitti_fast"]
    cam = synth.camera(cfg.camera)
    distinct, total = 8, 4096
    left, right = synth.band_world_batch(cfg.camera, range(300, 300 + distinct))
    idx = np.arange(total) % distinct
    big_l, big_r = np.ascontiguousarray(left[idx]), np.ascontiguousarray(right[idx])
    gen = api.StereoFramePointGenerator(cfg, cam, max_batch=total)
    out, counts = gen.batch_process(big_l, big_r, True)
    out2, nf, nm, nl, nr = gen.batch_download(total)
    assert np.array_equal(counts, nf)
    for i in range(distinct):
        o = _oracle(cfg, cam, left[i], right[i], True)
        _same_points(out[i, :counts[i]], o.framepoints())
        assert nm[i] == len(o.matches) and nl[i] == len(o.kps_left) and nr[i] == len(o.kps_right)
        k, d = gen.features(0, pair=total - distinct + i)
        assert np.array_equal(d, o.desc_left) and np.array_equal(k["x"], o.kps_left["x"])
    view = out.reshape(total // distinct, distinct, -1)
    for name, a in (("counts", counts), ("matches", nm), ("left", nl), ("right", nr)):
        assert np.array_equal(a.reshape(-1, distinct), np.broadcast_to(a[:distinct], (total // distinct, distinct))), name
    for i in range(distinct):
        ref = view[0, i, :counts[i]].tobytes()
        for r in range(1, total // distinct):
            assert view[r, i, :counts[i]].tobytes() == ref, (r, i)
    gen.close()


def _stress_images(kind, rows, cols, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:rows, 0:cols]
    if kind == "noise":                       # almost every pixel passes the pre-test; dense corners
        img = rng.integers(0, 256, (rows, cols))
    elif kind == "squares":                   # isolated bright 7 x 7 squares on black: L-corners on a lattice, almost equal
        img = (((yy % 13) < 7) & ((xx % 13) < 7)) * 249 + rng.integers(0, 7, (rows, cols))   # near-ties in the 3x3 NMS
    elif kind == "saturated_rectangles":      # overlapping flat 0 / 255 / mid-grey rectangles
        img = np.full((rows, cols), 128)
        for _ in range(rows * cols // 400):
            y, x, h, w = rng.integers(0, rows), rng.integers(0, cols), rng.integers(4, 30), rng.integers(4, 30)
            img[y:y + h, x:x + w] = rng.choice([0, 255, 64, 200])
    elif kind == "gradient_texture":          # smooth ramp + mid-frequency texture + sparse salt
        img = (xx * 255.0 / cols + 40 * np.sin(yy / 3.0) * np.sin(xx / 4.0)).astype(np.int64)
        salt = rng.random((rows, cols)) < 0.01
        img = np.where(salt, rng.integers(0, 256, (rows, cols)), img)
    else:
        raise ValueError(kind)
    left = np.clip(img, 0, 255).astype(np.uint8)
    right = np.roll(left, -6, axis=1)         # a constant 6 px disparity: plenty of true matches
    return np.ascontiguousarray(left), np.ascontiguousarray(right)


@pytest.mark.parametrize("kind", ["noise", "squares", "saturated_rectangles", "gradient_texture"])
@pytest.mark.parametrize("rows,cols,grid", [(217, 333, (1, 1)), (250, 515, (2, 2)), (131, 129, (1, 3))])
def test_stress_patterns_at_odd_sizes_match_oracle(kind, rows, cols, grid):
    """image statistics and shapes the band world does not produce: dense noise (every candidate list full), saturated
    lattices (blur rounding ties, equal scores in the 3x3 NMS), widths that are no multiple of 4 / 32 / 128, tiles cut
    by the image border in both directions, 1x3 and 2x2 detector grids.  Everything stays bit-exact."""
    import dataclasses
    cfg = dataclasses.replace(configs.EUROC, number_of_detectors_vertical=grid[0], number_of_detectors_horizontal=grid[1],
                              detector_threshold_minimum=7, detector_threshold_maximum=60)
    cam = synth.Camera(cols, rows, 300.0, 300.0, cols / 2.0, rows / 2.0, -30.0)
    left, right = _stress_images(kind, rows, cols, rows + cols)
    gen = api.StereoFramePointGenerator(cfg, cam, max_keypoints=40000)
    o = pipeline.StereoFramePointGeneratorOracle(cfg, cam, "a")
    for frame in range(2):                    # the second frame runs at the thresholds the controller proposed
        nl, nr = gen.initialize(left, right, frame == 0)
        o.initialize(left, right, frame == 0)
        assert (nl, nr) == (len(o.kps_left), len(o.kps_right))
        _check_pair(gen, o)
        assert np.array_equal(gen.debug_blurred(0), tier_a.gauss7_u8(left))
        fps = gen.compute()
        o.compute(None)
        assert gen.number_of_matches == len(o.matches)
        _same_points(fps, o.framepoints())
        assert np.array_equal(gen.thresholds, o.thresholds)
        left, right = np.ascontiguousarray(left[:, ::-1]), np.ascontiguousarray(right[:, ::-1])   # new content next frame
    gen.close()
