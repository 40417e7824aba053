"""Pins tier A (oracle/c/vslam_oracle.c, the bit-exact spec of the CUDA kernels) to THE REFERENCE ITSELF: oracle/_ref is the
reference's own, unmodified translation units (stereo_framepoint_generator.cpp, base_framepoint_generator.cpp,
intensity_feature_matcher.cpp, stereouv_aligner.cpp, uvd_aligner.cpp, frame.cpp, frame_point.cpp, landmark.cpp,
world_map.cpp, parameters.cpp, pose_tracker_3d.cpp) compiled from /root/reference against the functional third-party
stand-ins of oracle/shims (oracle/Makefile `_ref`).  What is compared here is everything the reference OWNS on the hot
path -- detector grid, threshold controller, triangulation distance, row scan with cursor, epipolar passes, bin rule
and output order, getPointInLeftCamera, track() with its order-dependent feature consumption, recoverPoints(), the
aligners' initialize / linearize / oneRound / converge, Landmark::update, the YAML parsing quirks.  The three OpenCV
primitives below it (FAST, ORB, Hamming) are pinned to OpenCV 4.13 by tests/test_oracle_vs_cv2.py and tests/golden/.

Tolerances: integers, indices, orders, descriptors, coordinates, chi-square errors and inlier flags: bit-exact.
H, b, total error: 1e-10 relative (the reference sums Eigen expressions, tier A scalars).  Converged pose: 1e-9.
"""
import os

import numpy as np
import pytest

from oracle import pipeline, ref, tier_a
from vslam_b200 import configs, synth

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref needs /root/reference or a prebuilt library")

YAML = {"kitti": "configuration_kitti.yaml", "kitti_fast": "configuration_kitti_fast.yaml",
        "euroc": "configuration_euroc.yaml"}
HAVE_YAML = os.path.isdir(os.path.join(ref.REFERENCE_ROOT, "configurations"))


def _session(name, **overrides):
    cfg = configs.BY_NAME[name]
    cam = synth.camera(cfg.camera)
    if HAVE_YAML:
        return cfg, cam, ref.Session(cam, YAML[name], **overrides)
    # no YAML files on this machine (GPU box): struct defaults + the effective values of configs.py
    a = configs.ALIGNER_BY_NAME[name]
    values = dict(
        target_number_of_keypoints_tolerance=cfg.target_number_of_keypoints_tolerance,
        detector_threshold_minimum=cfg.detector_threshold_minimum, detector_threshold_maximum=cfg.detector_threshold_maximum,
        detector_threshold_maximum_change=cfg.detector_threshold_maximum_change,
        number_of_detectors_vertical=cfg.number_of_detectors_vertical,
        number_of_detectors_horizontal=cfg.number_of_detectors_horizontal,
        maximum_reliable_depth_meters=cfg.maximum_reliable_depth_meters, bin_size_pixels=cfg.bin_size_pixels,
        enable_keypoint_binning=int(cfg.enable_keypoint_binning),
        maximum_matching_distance_triangulation=cfg.maximum_matching_distance_triangulation,
        minimum_disparity_pixels=cfg.minimum_disparity_pixels,
        maximum_epipolar_search_offset_pixels=cfg.maximum_epipolar_search_offset_pixels,
        error_delta_for_convergence=a.error_delta_for_convergence, maximum_error_kernel=a.maximum_error_kernel,
        damping=a.damping, maximum_number_of_iterations=a.maximum_number_of_iterations,
        minimum_number_of_inliers=a.minimum_number_of_inliers)
    values.update(overrides)
    return cfg, cam, ref.Session(cam, None, **values)


# ---- parameters ------------------------------------------------------------------------------------------------------
@pytest.mark.skipif(not HAVE_YAML, reason="configurations/*.yaml live in /root/reference")
@pytest.mark.parametrize("name", ["kitti", "kitti_fast", "euroc"])
def test_effective_parameters_of_the_references_own_parser(name):
    """vslam_b200.configs == what ParameterCollection::parseFromFile leaves in the structs, quirks included"""
    cfg, cam, s = _session(name)
    p = s.parameters()
    for key in ("target_number_of_keypoints_tolerance", "detector_threshold_minimum", "detector_threshold_maximum",
                "detector_threshold_maximum_change", "number_of_detectors_vertical", "number_of_detectors_horizontal",
                "maximum_reliable_depth_meters", "bin_size_pixels", "maximum_matching_distance_triangulation",
                "minimum_disparity_pixels", "maximum_epipolar_search_offset_pixels"):
        assert getattr(p, key) == getattr(cfg, key), key
    assert bool(p.enable_keypoint_binning) == cfg.enable_keypoint_binning
    a = configs.ALIGNER_BY_NAME[name]
    for key in ("error_delta_for_convergence", "maximum_error_kernel", "damping", "maximum_number_of_iterations",
                "minimum_number_of_inliers"):
        assert getattr(p, key) == getattr(a, key), key
    assert a.minimum_reliable_depth_meters == p.minimum_depth_meters                 # slam_assembly.cpp:69-70
    assert a.maximum_reliable_depth_meters == p.maximum_reliable_depth_meters
    assert p.detector_type == b"FAST"          # parsed in RGB_DEPTH mode only (parameters.cpp:341)
    assert p.use_matches == 1                  # struct default: the dead FLANN block runs in every stereo YAML
    if name == "kitti":                        # 51.2 does not parse as int32_t (parameters.cpp:323): the default stays
        assert p.maximum_matching_distance_triangulation == 0.2 * 256
    s.configure()                              # BRIEF / BRIEF-256 / ORB-256 all end in cv::ORB::create() (:187-224)
    assert s.parameters().descriptor_type == b"ORB"
    s.close()


@pytest.mark.skipif(not HAVE_YAML, reason="configurations/*.yaml live in /root/reference")
@pytest.mark.parametrize("name", ["kitti", "kitti_fast", "euroc"])
def test_committed_effective_values_are_what_the_yaml_parses_to(name):
    cam = synth.camera(configs.BY_NAME[name].camera)
    s = ref.Session(cam, YAML[name])
    p = s.parameters()
    for key, value in ref.effective_values(name).items():
        got = getattr(p, key)
        assert (got.decode() if isinstance(got, bytes) else got) == value, key
    s.close()


# ---- first frames: initialize() + compute() ------------------------------------------------------------------------------
def _assert_features(s, o):
    for side, (kps, desc) in enumerate(((o.kps_left, o.desc_left), (o.kps_right, o.desc_right))):
        xyr, d = s.features(side)
        assert np.array_equal(xyr[:, 0], kps["x"]) and np.array_equal(xyr[:, 1], kps["y"])
        assert np.array_equal(xyr[:, 2], kps["response"])
        assert np.array_equal(d, desc)
    assert np.array_equal(s.thresholds(), o.thresholds)
    assert s.triangulation_distance() == o.max_distance
    assert s.target_number_of_keypoints() == o.target_number_of_keypoints


def _assert_new_points(pts, fp, o):
    """frame->points() entries created by compute() against tier A's winners"""
    assert len(pts) == len(fp)
    for k in ("xl", "yl", "xr", "yr"):
        assert np.array_equal(pts[k], fp[k]), k
    assert np.array_equal(pts["distance"], fp["distance"].astype(np.float64))
    assert np.array_equal(pts["epipolar_offset"], fp["epipolar_offset"])
    assert np.array_equal(pts["cam"], fp["cam"])                                   # getPointInLeftCamera: bit-exact
    assert np.array_equal(pts["row"], fp["yl"].astype(np.int32)) and np.array_equal(pts["col"], fp["xl"].astype(np.int32))
    assert np.array_equal(pts["disparity"], (fp["xl"] - fp["xr"]).astype(np.float64))
    assert np.array_equal(pts["desc_left"], o.desc_left[fp["index_left"]])
    assert np.array_equal(pts["desc_right"], o.desc_right[fp["index_right"]])
    assert np.all(pts["index_previous"] == -1)


@pytest.mark.parametrize("name,seed,tracking", [("kitti", 5, False), ("kitti", 6, True), ("kitti_fast", 7, False),
                                                ("euroc", 8, False), ("euroc", 9, True)])
def test_first_frame_matches_the_reference(name, seed, tracking):
    cfg, cam, s = _session(name)
    s.configure()
    left, right = synth.band_world_pair(cfg.camera, seed)
    s.initialize(left, right, tracking=tracking)
    o = pipeline.StereoFramePointGeneratorOracle(cfg, cam, "a")
    o.initialize(left, right, not tracking)
    _assert_features(s, o)
    n = s.compute()
    o.compute()
    assert n == len(o.framepoints()) > 200
    _assert_new_points(s.points(), o.framepoints(), o)
    # the scan removed the matched features from both pools (stereo_framepoint_generator.cpp:417-424)
    assert len(s.remaining(0)) == len(o.kps_left) - len(o.matches)
    assert len(s.remaining(1)) == len(o.kps_right) - len(o.matches)
    s.close()


@pytest.mark.parametrize("name,overrides", [
    ("kitti_fast", dict(maximum_epipolar_search_offset_pixels=3)),          # executables/test_stereo_frontend.cpp
    ("kitti", dict(enable_keypoint_binning=0)),
    ("euroc", dict(maximum_epipolar_search_offset_pixels=1, minimum_disparity_pixels=8.0)),
    ("kitti", dict(maximum_matching_distance_triangulation=20.0, bin_size_pixels=40)),
    ("kitti_fast", dict(number_of_detectors_vertical=2, number_of_detectors_horizontal=3))])
def test_first_frame_variants(name, overrides):
    import dataclasses
    cfg, cam, s = _session(name, **overrides)
    s.configure()
    cfg = dataclasses.replace(cfg, **{k: (bool(v) if k == "enable_keypoint_binning" else v) for k, v in overrides.items()})
    rng = np.random.default_rng(3)
    left, right = synth.band_world_pair(cfg.camera, 21)
    if overrides.get("maximum_epipolar_search_offset_pixels"):
        right = np.roll(right, 1, axis=0)         # a vertical misalignment gives the later passes something to find
    s.initialize(left, right)
    o = pipeline.StereoFramePointGeneratorOracle(cfg, cam, "a")
    o.initialize(left, right, True)
    _assert_features(s, o)
    assert s.compute() == len(o.compute()["winners"]) > 50
    _assert_new_points(s.points(), o.framepoints(), o)
    if overrides.get("maximum_epipolar_search_offset_pixels"):
        assert np.any(s.points()["epipolar_offset"] != 0)
    del rng
    s.close()


# ---- sequences: track() -> recoverPoints() -> compute() ------------------------------------------------------------------
def _previous_points(pts):
    """the reference's own previous frame->points() as tier A's track() / recoverPoints() take them"""
    p = np.zeros(len(pts), tier_a.PREVIOUS_POINT)
    p["cam"], p["world"] = pts["cam"], pts["landmark_world"]
    p["desc_left"], p["desc_right"] = pts["desc_left"], pts["desc_right"]
    p["epipolar_offset"], p["has_landmark"] = pts["epipolar_offset"], pts["has_landmark"]
    p["keypoint_size"] = 7.0
    return p


def _motion(cam, noise, seed):
    T = np.hstack([np.eye(3), np.zeros((3, 1))])
    T[0, 3] = -(-cam.bx / cam.fx) / 4            # previous -> current for the band world's B/4 step
    if noise:
        rng = np.random.default_rng(seed)
        T[:, 3] += rng.normal(0, noise, 3)
        T[:, :3] = synth._rot(*rng.normal(0, noise * 0.02, 3))
    return T


@pytest.mark.parametrize("name,by_appearance,distance,noise,overrides", [
    ("kitti_fast", False, 15, 0.02, {}), ("kitti", True, 50, 0.0, {}), ("euroc", False, 25, 0.01, {}),
    ("kitti", False, 20, 0.03, dict(maximum_epipolar_search_offset_pixels=2))])
def test_tracked_sequence_matches_the_reference(name, by_appearance, distance, noise, overrides):
    import dataclasses
    cfg, cam, s = _session(name, **overrides)
    s.configure()
    cfg = dataclasses.replace(cfg, **overrides)
    p = s.parameters()
    world = synth.BandWorld(cam.cols, cam.rows, 31, max_frames=8)
    o = pipeline.StereoFramePointGeneratorOracle(cfg, cam, "a")
    pose = np.hstack([np.eye(3), np.zeros((3, 1))])
    for k in range(5):
        left, right = world.pair(k)
        tracking = k >= 2
        s.initialize(left, right, tracking=tracking)
        o.initialize(left, right, not tracking)
        _assert_features(s, o)
        n_tracked = 0
        if k:
            prev = s.points(previous=True)
            T = _motion(cam, noise, seed=k)
            max_distance = 25.6 + 6.4 * k
            s.set_tracking(distance, max_distance)
            r = s.track(T, by_appearance)
            w = o.track(_previous_points(prev), T, by_appearance, distance, max_distance)
            t = w["tracks"]
            pts = s.points()
            n_tracked = len(pts)
            assert n_tracked == r["n_tracks"] == len(t) > 50
            assert np.array_equal(pts["index_previous"], t["index_previous"])
            for key in ("xl", "yl", "xr", "yr", "epipolar_offset", "cam"):
                assert np.array_equal(pts[key], t[key]), key
            assert np.array_equal(pts["distance"], t["distance"].astype(np.float64))
            for key in ("projection_left", "projection_right", "projection_right_corrected"):
                assert np.array_equal(pts[key], t[key]), key
            assert np.array_equal(r["lost"], w["lost"])
            assert r["number_of_tracked_landmarks"] == w["tracked_landmarks"]
            assert r["average_descriptor_distance"] == w["accumulated_distance"] / len(t)
            # prune(): the pools compute() scans afterwards
            for side, feats in ((0, o.features_left), (1, o.features_right)):
                rem = s.remaining(side)          # still in detection (region-major) order: compute() sorts (:159-160)
                rem = rem[np.lexsort((rem[:, 0], rem[:, 1]))]
                assert np.array_equal(rem[:, 0], feats["x"]) and np.array_equal(rem[:, 1], feats["y"])
            # recoverPoints() with the frame's pose set (pose_tracker_3d.cpp:170-191)
            pose = pose.copy()
            pose[0, 3] += (-cam.bx / cam.fx) / 4
            s.set_pose(pose)
            world_to_camera = np.hstack([pose[:, :3].T, -pose[:, :3].T @ pose[:, 3:]])
            n_recovered = s.recover()
            rec = o.recover_points(_previous_points(prev)[w["lost"]], world_to_camera, max_distance,
                                   p.minimum_depth_meters, p.maximum_depth_meters)
            assert n_recovered == len(rec)
            if k >= 3:
                assert n_recovered > 0
            pts = s.points()[n_tracked:]
            assert np.array_equal(pts["index_previous"], w["lost"][rec["index_lost"]])
            for key in ("xl", "yl", "xr", "yr", "cam", "desc_left", "desc_right"):
                assert np.array_equal(pts[key], rec[key]), key
            assert np.array_equal(pts["distance"], rec["distance"].astype(np.float64))
            n_tracked += n_recovered
        # compute(): tracked (and recovered) points pre-load the bins (:147-155)
        pts = s.points()
        tracked = np.zeros(len(pts), tier_a.TRACKED)
        tracked["row"], tracked["col"], tracked["has_previous"] = pts["row"], pts["col"], 1
        tracked["disparity"], tracked["distance"] = pts["disparity"], pts["distance"]
        n = s.compute()
        o.compute(tracked if len(tracked) else None)
        new = o.framepoints()
        assert n == n_tracked + len(new)
        _assert_new_points(s.points()[n_tracked:], new, o)
        # every second tracked point becomes / updates a landmark, as PoseTracker3D::_updatePoints would (:486-519)
        if k:
            s.make_landmarks(2)
    s.close()


# ---- aligners ---------------------------------------------------------------------------------------------------------------
def _problem(kind, n, cam, seed):
    c = synth.correspondences(n, "stereouv" if kind == 0 else "uvd", cam, seed=seed)
    if kind == 1:
        c["omega"] = np.ascontiguousarray(np.stack([c["omega_uv"], c["omega_d"]], 1))
    return c


def _close(a, b, tol=1e-10):
    scale = max(np.abs(b).max(), 1e-300)
    return np.abs(np.asarray(a) - np.asarray(b)).max() <= tol * scale


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("name", ["kitti", "kitti_fast"])
def test_linearize_matches_the_reference(kind, name):
    cfg, cam, s = _session(name)
    s.configure()
    a = configs.ALIGNER_BY_NAME[name]
    c = _problem(kind, 5000, cam, seed=99 + kind)
    K = np.array([[cam.fx, 0, cam.cx], [0, cam.fy, cam.cy], [0, 0, 1.0]])
    baseline = np.array([cam.bx, 0.0, 0.0])
    s.aligner_load(kind, c["moving"], c["fixed"], c["omega"], c["wt"], baseline, a.minimum_reliable_depth_meters)
    A = tier_a.Aligner(kind, c["moving"], c["fixed"], c["omega"], c["wt"], K, baseline, cam.rows, cam.cols,
                       a.minimum_reliable_depth_meters, a.maximum_error_kernel)
    poses = (np.hstack([np.eye(3), np.zeros((3, 1))]), synth.true_motion(0.7)[:3], synth.true_motion(1.0)[:3])
    for T in poses:
        for ignore in (False, True):
            s.aligner_set_pose(kind, T)
            r = s.aligner_linearize(kind, ignore)
            w = A.linearize(T, ignore)
            assert np.array_equal(r["errors"], A.errors)                            # chi-square per point: bit-exact
            assert np.array_equal(r["inlier_flags"], A.inliers.astype(bool))
            assert r["inliers"] == w["inliers"] and r["outliers"] == w["outliers"]
            if T is poses[-1]:
                assert 0.3 * 5000 < r["inliers"] < 5000
            assert abs(r["total_error"] - w["total_error"]) <= 1e-10 * w["total_error"]
            assert _close(r["H"], w["H"]) and np.abs(r["b"] - w["b"]).max() <= 1e-10 * np.abs(w["b"]).max() + \
                1e-13 * np.sqrt(np.abs(np.diag(w["H"])).max() * w["total_error"])
    s.close()


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("name", ["kitti", "kitti_fast", "euroc"])
def test_converge_matches_the_reference(kind, name):
    cfg, cam, s = _session(name)
    s.configure()
    a = configs.ALIGNER_BY_NAME[name]
    c = _problem(kind, 3000, cam, seed=7 + kind)
    K = np.array([[cam.fx, 0, cam.cx], [0, cam.fy, cam.cy], [0, 0, 1.0]])
    baseline = np.array([cam.bx, 0.0, 0.0])
    s.aligner_load(kind, c["moving"], c["fixed"], c["omega"], c["wt"], baseline, a.minimum_reliable_depth_meters)
    A = tier_a.Aligner(kind, c["moving"], c["fixed"], c["omega"], c["wt"], K, baseline, cam.rows, cam.cols,
                       a.minimum_reliable_depth_meters, a.maximum_error_kernel)
    # one round first: damping, FullPivLU solve, v2t, re-orthonormalisation
    T0 = np.hstack([np.eye(3), np.zeros((3, 1))])
    r1 = s.aligner_one_round(kind, False)
    T1, _ = A.one_round(T0, a.damping, False)
    assert np.abs(r1["T"] - T1).max() < 1e-12
    s.aligner_set_pose(kind, T0)
    r = s.aligner_converge(kind)
    # UVDAligner::converge hard-codes `inliers > 100` where StereoUV reads minimum_number_of_inliers (uvd_aligner.cpp:207)
    w = A.converge(T0, a.damping, a.error_delta_for_convergence, a.maximum_number_of_iterations,
                   a.minimum_number_of_inliers)
    assert r["converged"] == bool(w["converged"]) is True
    assert r["rounds"] - 1 == w["rounds"]                      # - the oneRound above
    assert np.abs(r["T"] - w["T"]).max() < 1e-9
    assert r["inliers"] == w["inliers"] and np.array_equal(r["inlier_flags"], A.inliers.astype(bool))
    assert _close(r["information"], w["information"], 1e-9)
    truth = synth.true_motion(1.0)[:3]
    assert np.abs(r["T"] - truth).max() < 0.05
    s.close()


def test_aligner_initialize_on_the_references_frames():
    """StereoUVAligner::initialize (stereouv_aligner.cpp:10-69) on two frames the reference's own generator produced: the
    packed arrays are what the adapters upload (adapters/gpu_frame_aligners.cpp) and what linearize_pairs_kernel reads"""
    cfg, cam, s = _session("kitti")
    s.configure()
    world = synth.BandWorld(cam.cols, cam.rows, 12, max_frames=4)
    T = _motion(cam, 0.0, 0)
    stale = np.zeros(0)
    for k in range(3):
        s.initialize(*world.pair(k), tracking=k >= 2)
        if k:
            s.set_tracking(15, 38.4)
            s.track(T, False)
            s.make_landmarks(2)
            pts, prev = s.points(), s.points(previous=True)
            for inverse_depth in (False, True):
                n = s.aligner_initialize_frames(0, T, inverse_depth)
                moving, fixed, omega, wt = s.aligner_packed(0)
                assert n == len(pts) > 300
                assert np.array_equal(fixed, np.stack([pts["xl"], pts["yl"], pts["xr"], pts["yr"]], 1).astype(np.float64))
                q = prev[pts["index_previous"]]
                has = q["has_landmark"].astype(bool)
                assert np.array_equal(moving[~has], q["cam"][~has])
                if k == 2:
                    assert has.any()
                    assert np.array_equal(omega[has], 1 + np.log(q["landmark_updates"][has]))
                assert np.all(omega[~has] == 1.0)
                # quirk kept: `_weights_translation.resize(n, 1)` (:24) only initialises NEW entries, so with
                # enable_inverse_depth_as_information off (Localizing, pose_tracker_3d.cpp:124) the weights of the last
                # call survive in the first min(n, n_last) slots
                if inverse_depth:
                    want = np.minimum(15.0 / pts["cam"][:, 2], 1.0)
                else:
                    want = np.ones(n)
                    want[:min(n, len(stale))] = stale[:n]
                assert np.array_equal(wt, want)
                stale = wt.copy()
            r = s.aligner_converge(0)
            assert r["converged"] and r["inliers"] > 0.3 * n
            # the same problem through tier A: round count, pose, inlier set
            a = configs.ALIGNER_BY_NAME["kitti"]
            K = np.array([[cam.fx, 0, cam.cx], [0, cam.fy, cam.cy], [0, 0, 1.0]])
            A = tier_a.Aligner(0, moving, fixed, omega, wt, K, [cam.bx, 0.0, 0.0], cam.rows, cam.cols,
                               a.minimum_reliable_depth_meters, a.maximum_error_kernel)
            w = A.converge(T, a.damping, a.error_delta_for_convergence, a.maximum_number_of_iterations,
                           a.minimum_number_of_inliers)
            assert w["converged"] and w["rounds"] == r["rounds"]
            assert np.abs(r["T"] - w["T"]).max() < 1e-9
            assert np.array_equal(r["inlier_flags"], A.inliers.astype(bool))
        s.compute()
    s.close()


# ---- Landmark::update --------------------------------------------------------------------------------------------------
def test_landmark_update_matches_the_reference():
    cfg, cam, s = _session("kitti")
    s.configure()
    p = s.parameters()
    h = synth.landmark_histories(40, n_frames=12, seed=3)
    poses_rw = h["camera_to_world"].reshape(-1, 3, 4)
    checked = 0
    for i in range(40):
        lo, hi = h["offsets"][i], h["offsets"][i + 1]
        m = h["measurements"][lo:hi]
        if len(m) < 3:
            continue
        world_ref, updates_ref = s.landmark_run(m["frame"], m["camera_coordinates"], poses_rw)
        # tier A: the constructor's average over the first two points (landmark.cpp:20-33), then one update per point
        first = [poses_rw[m["frame"][j]][:, :3] @ m["camera_coordinates"][j] + poses_rw[m["frame"][j]][:, 3] for j in (1, 0)]
        world = (first[0] + first[1]) / 2
        updates = 2
        order = [1, 0] + list(range(2, len(m)))       # the constructor walks the track backwards
        for n in range(3, len(m) + 1):
            mm = m[order[:n]].copy()
            world, updates, outcome, _ = tier_a.landmark_update(
                mm, h["world_to_camera"], h["camera_to_world"], world, updates,
                maximum_error_squared_meters=p.maximum_error_squared_meters)
        assert updates_ref == updates
        assert np.abs(world_ref - world).max() <= 1e-12 * max(1.0, np.abs(world).max())
        checked += 1
    assert checked > 20
    s.close()
