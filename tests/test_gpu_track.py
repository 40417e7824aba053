"""Parity of the device-side StereoFramePointGenerator::track / recoverPoints (through the C ABI) with the CPU oracle:
bit-exact tracks (previous index, feature indices, distance, epipolar offset, projections, triangulated point), lost
lists, landmark counts, average distance, the pruned candidate pools (checked through the following compute()), and
recovered points with their descriptors.  The order-dependent consumption of features (a tracked point removes its
features for every later point) is stressed with duplicated previous points."""
import numpy as np
import pytest

from oracle import pipeline, tier_a
from vslam_b200 import api, configs, synth

from test_oracle_track import _motion, previous_points

pytestmark = pytest.mark.gpu

TRACK_FIELDS = ("index_previous", "index_left", "index_right", "xl", "yl", "xr", "yr", "distance", "epipolar_offset",
                "projection_left", "projection_right", "projection_right_corrected")


def _as_api(prev):
    return prev.view(api.PREVIOUS_POINT)


def _same_tracks(got, want):
    assert len(got["tracks"]) == len(want["tracks"])
    for f in TRACK_FIELDS:
        assert np.array_equal(got["tracks"][f], want["tracks"][f]), f
    assert np.array_equal(got["tracks"]["camera"], want["tracks"]["cam"])       # bit-exact (stated bound: 1e-4 rel)
    assert np.array_equal(got["lost"], want["lost"])
    assert got["tracked_landmarks"] == want["tracked_landmarks"]
    n = len(want["tracks"])
    if n:
        assert got["average_descriptor_distance"] == want["accumulated_distance"] / n
    else:
        assert np.isnan(got["average_descriptor_distance"])


def _same_points(got, want):
    assert len(got) == len(want)
    for f in ("index_left", "index_right", "xl", "yl", "xr", "yr", "distance", "epipolar_offset"):
        assert np.array_equal(got[f], want[f]), f
    assert np.array_equal(got["camera"], want["cam"])


def _setup(cfg, seed, k0=0, k1=1, max_keypoints=0):
    cam = synth.camera(cfg.camera)
    world = synth.BandWorld(cam.cols, cam.rows, seed, max_frames=8)
    ora = pipeline.StereoFramePointGeneratorOracle(cfg, cam, "a")
    gen = api.StereoFramePointGenerator(cfg, cam, max_keypoints=max_keypoints)
    l0, r0 = world.pair(k0)
    ora.initialize(l0, r0, True)
    ora.compute()
    prev = previous_points(ora, ora.framepoints())
    l1, r1 = world.pair(k1)
    gen.initialize(l0, r0, True)      # the generator's thresholds follow the same two-frame history
    gen.compute()
    ora.initialize(l1, r1, False)
    gen.initialize(l1, r1, False)
    return cam, world, ora, gen, prev


@pytest.mark.parametrize("cfg_name,by_appearance,D,noise,dup", [
    ("kitti_fast", True, 50, 0.0, 0), ("kitti_fast", False, 15, 0.02, 7), ("euroc", True, 50, 0.01, 5),
    ("euroc", False, 25, 0.0, 0), ("kitti", False, 15, 0.05, 3), ("kitti", True, 50, 0.0, 1), ("hd", False, 30, 0.01, 4)])
def test_track_and_following_compute_match_oracle(cfg_name, by_appearance, D, noise, dup):
    cfg = configs.BY_NAME[cfg_name]
    cam, world, ora, gen, prev = _setup(cfg, seed=11)
    T = _motion(cam, 1, noise, seed=5)
    if dup:   # duplicates compete for the same features: the lower index must win, the others re-pick or get lost
        extra = prev[::dup].copy()
        extra["cam"][:, 0] += 0.01
        prev = np.concatenate([prev, extra, prev[::dup + 4]])
        prev["has_landmark"][::3] = 0
    want = ora.track(prev, T, by_appearance, D, 38.4)
    got = gen.track(_as_api(prev), T, by_appearance, D, 38.4)
    assert len(want["tracks"]) > 100
    _same_tracks(got, want)
    # compute(): the pruned pools and the bin pre-load must agree as well
    ora.compute(ora.tracked_points(want["tracks"]))
    fps = gen.compute(api.TRACKED_FROM_LAST_TRACK)
    assert gen.number_of_matches == len(ora.matches)
    _same_points(fps, ora.framepoints())
    # ... and the host-supplied pre-load gives the same result as the device-resident one
    gen.thresholds = ora.thresholds_used       # the frame is processed again with the thresholds it was detected with
    gen.initialize(*world.pair(1), False)
    got2 = gen.track(_as_api(prev), T, by_appearance, D, 38.4)
    assert np.array_equal(got2["tracks"], got["tracks"])
    tracked = np.zeros(len(got2["tracks"]), api.TRACKED)
    tracked["row"], tracked["col"] = got2["tracks"]["yl"].astype(np.int32), got2["tracks"]["xl"].astype(np.int32)
    tracked["has_previous"] = 1
    tracked["disparity"] = (got2["tracks"]["xl"] - got2["tracks"]["xr"]).astype(np.float64)
    tracked["distance"] = got2["tracks"]["distance"]
    assert np.array_equal(gen.compute(tracked), fps)
    gen.close()


@pytest.mark.parametrize("max_keypoints", [0, 4097, 40000])
def test_track_heavy_conflicts(max_keypoints):
    """every previous point three times, in shuffled order, with a loose appearance gate and the widest window: hundreds
    of points pick a feature a lower-indexed point already consumed.  With the default capacity the resolver keeps its
    claims, tentative results and worklist in shared memory (also with an odd capacity, which shifts the layout); a handle
    created for 40 000 keypoints per image does not fit (320 KB of claims) and runs the same rounds on the global scratch."""
    cfg = configs.KITTI_FAST
    cam, world, ora, gen, prev = _setup(cfg, seed=3, max_keypoints=max_keypoints)
    rng = np.random.default_rng(0)
    prev = np.concatenate([prev, prev, prev])[rng.permutation(3 * len(prev))]
    T = _motion(cam, 1, 0.03, seed=1)
    want = ora.track(prev, T, True, 50, 51.2)
    got = gen.track(_as_api(prev), T, True, 50, 51.2)
    _same_tracks(got, want)
    assert len(want["lost"]) > 100 and len(want["tracks"]) > 300
    ora.compute(ora.tracked_points(want["tracks"]))
    _same_points(gen.compute(api.TRACKED_FROM_LAST_TRACK), ora.framepoints())
    gen.close()


def test_track_degenerate_inputs():
    cfg = configs.KITTI_FAST
    cam, world, ora, gen, prev = _setup(cfg, seed=2)
    p = prev[:8].copy()
    p["cam"][0] = [0.0, 0.0, 0.0]              # 0/0 -> NaN
    p["cam"][1] = [1e9, 0.0, 1.0]              # far outside
    p["cam"][2] = [0.0, 0.0, -5.0]             # behind the camera
    p["cam"][4] = [1e300, 1e300, 1e-300]       # overflow to inf
    p["desc_left"][3] = ~p["desc_left"][3]
    T = _motion(cam)
    _same_tracks(gen.track(_as_api(p), T, True, 50, 38.4), ora.track(p, T, True, 50, 38.4))
    gen.initialize(*world.pair(1), False)
    ora.initialize(*world.pair(1), False)
    empty = gen.track(_as_api(prev[:0]), T, True, 50, 38.4)
    assert len(empty["tracks"]) == 0 and len(empty["lost"]) == 0 and np.isnan(empty["average_descriptor_distance"])
    ora.compute()
    _same_points(gen.compute(api.TRACKED_FROM_LAST_TRACK), ora.framepoints())
    # zero-pixel window and a window that covers the whole image
    for D in (0, 2000):
        gen.initialize(*world.pair(1), False)
        ora.initialize(*world.pair(1), False)
        _same_tracks(gen.track(_as_api(prev), T, False, D, 38.4), ora.track(prev, T, False, D, 38.4))
    gen.close()


def test_compute_from_last_track_requires_track():
    cfg = configs.KITTI_FAST
    cam = synth.camera(cfg.camera)
    gen = api.StereoFramePointGenerator(cfg, cam)
    gen.initialize(*synth.band_world_pair(cfg.camera, 0), True)
    with pytest.raises(api.VslamError):
        gen.compute(api.TRACKED_FROM_LAST_TRACK)
    gen.close()


@pytest.mark.parametrize("cfg_name", ["euroc", "kitti"])
def test_tracked_sequence_matches_oracle(cfg_name):
    """initialize -> track -> compute over consecutive frames, the previous points of frame k+1 being ALL points of
    frame k (tracks first, then the new points, like frame->points()), built from the generator's own outputs"""
    cfg = configs.BY_NAME[cfg_name]
    cam = synth.camera(cfg.camera)
    world = synth.BandWorld(cam.cols, cam.rows, 21, max_frames=8)
    ora = pipeline.StereoFramePointGeneratorOracle(cfg, cam, "a")
    gen = api.StereoFramePointGenerator(cfg, cam)
    T = _motion(cam)
    prev_o = prev_g = None
    D, dist = 50, 25.6
    for k in range(5):
        left, right = world.pair(k)
        ora.initialize(left, right, k == 0)
        gen.initialize(left, right, k == 0)
        kl, dl = gen.features(0)
        kr, dr = gen.features(1)
        if prev_o is None:
            ora.compute()
            fps = gen.compute()
            _same_points(fps, ora.framepoints())
            pts_o = [ora.framepoints()]
            pts_g = [fps]
        else:
            want = ora.track(prev_o, T, k < 2, D, dist)
            got = gen.track(prev_g, T, k < 2, D, dist)
            _same_tracks(got, want)
            assert len(want["tracks"]) > 0.5 * len(prev_o)
            ora.compute(ora.tracked_points(want["tracks"]))
            fps = gen.compute(api.TRACKED_FROM_LAST_TRACK)
            _same_points(fps, ora.framepoints())
            pts_o = [want["tracks"], ora.framepoints()]
            pts_g = [got["tracks"], fps]
            D = max(15, D - 10)
        # frame->points() of this frame -> the previous points of the next one
        def build(parts, desc_l, desc_r, cam_field, dtype):
            n = sum(len(p) for p in parts)
            out = np.zeros(n, dtype)
            names = out.dtype.names
            i = 0
            for p in parts:
                s = slice(i, i + len(p))
                out[names[0]][s] = p[cam_field]
                out[names[1]][s] = p[cam_field]
                out[names[2]][s] = desc_l[p["index_left"]]
                out[names[3]][s] = desc_r[p["index_right"]]
                out["epipolar_offset"][s] = p["epipolar_offset"]
                i += len(p)
            out["has_landmark"] = 1
            out["keypoint_size"] = 7.0
            return out
        prev_o = build(pts_o, ora.desc_left, ora.desc_right, "cam", tier_a.PREVIOUS_POINT)
        prev_g = build(pts_g, dl, dr, "camera", api.PREVIOUS_POINT)
        assert prev_o.tobytes() == prev_g.tobytes()
    gen.close()


@pytest.mark.parametrize("cfg_name", ["kitti", "euroc"])
def test_recover_points_match_oracle(cfg_name):
    cfg = configs.BY_NAME[cfg_name]
    cam, world, ora, gen, prev = _setup(cfg, seed=9)
    lost = prev.copy()
    lost["has_landmark"][::5] = 0
    lost["keypoint_size"][1::9] = 5.0          # border 25 < 31: ORB would drop the keypoint
    lost["world"][2::11, 2] *= 400.0           # beyond maximum_depth_meters
    W = _motion(cam, 1, 0.01, seed=2)
    for max_track, max_depth in ((38.4, 1000.0), (64.0, 100.0), (-1.0, 1000.0)):
        want = ora.recover_points(lost, W, max_track, 0.1, max_depth)
        got = gen.recover_points(_as_api(lost), W, max_track, 0.1, max_depth)
        assert len(got) == len(want)
        for f, g in (("index_lost", "index_lost"), ("distance", "distance"), ("xl", "xl"), ("yl", "yl"), ("xr", "xr"),
                     ("yr", "yr"), ("camera", "cam"), ("descriptor_left", "desc_left"),
                     ("descriptor_right", "desc_right")):
            assert np.array_equal(got[f], want[g]), f
    assert len(ora.recover_points(lost, W, 38.4)) > 100
    assert len(gen.recover_points(_as_api(lost[:0]), W, 38.4)) == 0
    gen.close()


def test_initialize_without_extraction_forgets_the_abandoned_track():
    """PoseTracker3D retries a failed track() after initialize(frame, extract_features=false)
    (pose_tracker_3d.cpp:320, 402; stereo_framepoint_generator.cpp:126-133): both matchers are set up again from the
    frame's features, so the second track() sees every feature -- its result must equal a track() on a freshly
    initialized frame, whatever the abandoned attempt pruned; and a compute() after the reset scans all features."""
    cfg = configs.BY_NAME["kitti"]
    cam, world, ora, gen, prev = _setup(cfg, seed=11)
    T_bad = _motion(cam, 3, 0.0, seed=5)            # a poor prior: few tracks, but they prune features
    T_good = _motion(cam, 1, 0.0, seed=5)
    first = gen.track(_as_api(prev), T_bad, False, 15, 38.4)
    gen.reset_features()                             # initialize(current_frame_, false)
    second = gen.track(_as_api(prev), T_good, True, 50, 38.4)
    want = ora.track(prev, T_good, True, 50, 38.4)   # the oracle's frame was never tracked with the bad prior
    assert len(want["tracks"]) > 100 and len(first["tracks"]) != len(second["tracks"])
    _same_tracks(second, want)
    ora.compute(ora.tracked_points(want["tracks"]))
    fps = gen.compute(api.TRACKED_FROM_LAST_TRACK)
    assert gen.number_of_matches == len(ora.matches)
    _same_points(fps, ora.framepoints())
    # after a reset the tracks of the abandoned attempt cannot pre-load compute()
    gen.reset_features()
    with pytest.raises(api.VslamError):
        gen.compute(api.TRACKED_FROM_LAST_TRACK)
    gen.close()


def test_prune_tracks_on_the_device_equals_prune_on_the_host():
    """PoseTracker3D::_prunePoints (pose_tracker_3d.cpp:437-472): compute() after vslam_fpg_prune_tracks must give what
    compute() gives when the host prunes the tracks itself and uploads the survivors as tracked points -- both branches
    of the rule (inliers only / error cap), against the oracle's bin rule"""
    cfg, acfg = configs.KITTI, configs.KITTI_ALIGNER
    cam = synth.camera(cfg.camera)
    world = synth.BandWorld(cam.cols, cam.rows, 23, max_frames=4)
    for kernel, expect_inliers_only in ((acfg.maximum_error_kernel, False), (1e6, True)):
        gen = api.StereoFramePointGenerator(cfg, cam)
        ora = pipeline.StereoFramePointGeneratorOracle(cfg, cam, "a")
        gen.initialize(*world.pair(0), True)
        ora.initialize(*world.pair(0), True)
        first = gen.compute()
        ora.compute()
        _, dl = gen.features(0)
        _, dr = gen.features(1)
        prev = api.make_previous_points([first], dl, dr)
        gen.initialize(*world.pair(1), False)
        ora.initialize(*world.pair(1), False)
        T = np.hstack([np.eye(3), np.zeros((3, 1))])
        T[0, 3] = -(-cam.bx / cam.fx) / 4 + 0.05            # a poor prior: the aligner rejects a good share of the tracks
        got = gen.track(prev, T, False, 25, 40.0)
        ora.track(prev.view(tier_a.PREVIOUS_POINT), T, False, 25, 40.0)
        tr = got["tracks"]
        assert len(tr) > 200
        # a deliberately wrong pose and ONE linearisation: many outliers, errors spread around the cap
        import dataclasses
        al = api.StereoUVAligner(dataclasses.replace(acfg, maximum_error_kernel=kernel), max_points=4096)
        moving = np.ascontiguousarray(prev["camera_left"][tr["index_previous"]])
        fixed = np.stack([tr["xl"], tr["yl"], tr["xr"], tr["yr"]], 1).astype(np.float64)
        al.initialize(moving, fixed, np.ones(len(tr)), np.ones(len(tr)), cam.K, cam.baseline, cam.rows, cam.cols, T)
        sys_ = al.linearize(False)
        errors, inliers = al.errors(), al.inliers()
        average = sys_["total_error"] / len(tr)
        assert (average < kernel) == expect_inliers_only
        want_keep = inliers if average < kernel else (errors != -1) & (errors < 100 * kernel)
        keep = gen.prune_tracks(al, kernel)
        assert np.array_equal(keep, want_keep) and 0 < keep.sum()
        if not expect_inliers_only:
            assert keep.sum() < len(tr)
        new_device = gen.compute(api.TRACKED_FROM_LAST_TRACK)
        ora.compute(ora.tracked_points(ora.tracks[want_keep]))
        want = ora.framepoints()
        assert len(new_device) == len(want) and np.array_equal(new_device["index_left"], want["index_left"])
        assert np.array_equal(new_device["camera"], want["cam"])
        gen.close()
        al.close()
