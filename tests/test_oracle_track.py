"""Pins the C oracle of StereoFramePointGenerator::track / recoverPoints
(/root/reference/src/framepoint_generation/stereo_framepoint_generator.cpp:464-681, 683-869;
intensity_feature_matcher.cpp:81-148) against an independent, literal Python restatement of the same reference
lines (dict lattice, plain loops) on band-world sequences, and recoverPoints' descriptors against cv::ORB::compute on
the 71 x 71 region of interest exactly as the reference calls it."""
import numpy as np
import pytest

from oracle import pipeline, tier_a
from vslam_b200 import configs, synth


def previous_points(ora, fps, has_landmark=None, world=None):
    """frame->points() of a processed frame as track()/recoverPoints() read them"""
    p = np.zeros(len(fps), tier_a.PREVIOUS_POINT)
    p["cam"] = fps["cam"]
    p["world"] = fps["cam"] if world is None else world
    p["desc_left"] = ora.desc_left[fps["index_left"]]
    p["desc_right"] = ora.desc_right[fps["index_right"]]
    p["epipolar_offset"] = fps["epipolar_offset"]
    p["has_landmark"] = 1 if has_landmark is None else has_landmark
    p["keypoint_size"] = 7.0
    return p


def _i32(v):
    if not (v > -2147483649.0 and v < 2147483648.0):   # NaN / out of range: x86 "integer indefinite"
        return -2 ** 31
    return int(v)       # truncation toward zero


def _ham(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def _region(lattice, f, row_ref, col_ref, desc, r0, r1, c0, c1, max_dist, by_appearance):
    best, best_d = None, float(max_dist)
    if by_appearance:
        for row in range(r0, r1):
            for col in range(c0, c1):
                k = lattice.get((row, col))
                if k is None:
                    continue
                d = _ham(desc, f["desc"][k])
                if d < best_d:
                    best_d, best = d, k
    else:
        best_p = 10000
        for row in range(r0, r1):
            for col in range(c0, c1):
                k = lattice.get((row, col))
                if k is None:
                    continue
                d = _ham(desc, f["desc"][k])
                if d < max_dist:
                    pd = (row_ref - row) ** 2 + (col_ref - col) ** 2
                    if pd < best_p:
                        best_p, best_d, best = pd, d, k
    return best, best_d


def py_track(fl, fr, rows, cols, cam, prev, T, by_appearance, D, max_track, max_tri, min_disp):
    ll = {(int(f["row"]), int(f["col"])): i for i, f in enumerate(fl)}
    lr = {(int(f["row"]), int(f["col"])): i for i, f in enumerate(fr)}
    ml, mr = np.zeros(len(fl), bool), np.zeros(len(fr), bool)
    tracks, lost = [], []
    for u, pp in enumerate(prev):
        pc = [T[i, 0] * pp["cam"][0] + T[i, 1] * pp["cam"][1] + T[i, 2] * pp["cam"][2] + T[i, 3] for i in range(3)]
        il = [cam.fx * pc[0] + cam.cx * pc[2], cam.fy * pc[1] + cam.cy * pc[2], pc[2]]
        with np.errstate(all="ignore"):
            col_l, row_l = _i32(np.float64(il[0]) / np.float64(il[2])), _i32(np.float64(il[1]) / np.float64(il[2]))
        if col_l < 0 or col_l > cols or row_l < 0 or row_l > rows:
            continue
        kl, _ = _region(ll, fl, row_l, col_l, pp["desc_left"], max(row_l - D, 0), min(row_l + D + 1, rows),
                        max(col_l - D, 0), min(col_l + D + 1, cols), max_track, by_appearance)
        tracked = False
        if kl is not None:
            f_l = fl[kl]
            ex = np.float32(col_l) - f_l["x"]
            ey = np.float32(row_l) - f_l["y"]
            ir = [il[0] + cam.bx, il[1], il[2]]
            col_r, row_r = _i32(ir[0] / ir[2] - float(ex)), _i32(ir[1] / ir[2] - float(ey))
            if col_r < 0 or col_r > cols or row_r < 0 or row_r > rows:
                continue
            e = int(abs(float(pp["epipolar_offset"])))
            kr, dist = _region(lr, fr, row_r, col_r, f_l["desc"], max(row_r - e, 0), min(row_r + e + 1, rows),
                               max(col_r - D, 0), min(col_r + D + 1, int(f_l["col"])), max_tri, True)
            if kr is not None:
                f_r = fr[kr]
                if int(f_l["col"]) - int(f_r["col"]) < min_disp:
                    continue
                if _ham(f_r["desc"], pp["desc_right"]) > max_track:
                    continue
                for col in range(int(f_r["col"]) + 1, int(f_l["col"])):
                    k = lr.pop((int(f_r["row"]), col), None)
                    if k is not None:
                        mr[k] = True
                tracks.append((u, int(f_l["index"]), int(f_r["index"]), int(dist), int(f_r["row"]) - int(f_l["row"]),
                               col_l, row_l, col_r, row_r))
                ml[kl] = mr[kr] = True
                del ll[(int(f_l["row"]), int(f_l["col"]))]
                del lr[(int(f_r["row"]), int(f_r["col"]))]
                tracked = True
        if not tracked:
            lost.append(u)
    return tracks, lost, ml, mr


def _two_frames(cfg, seed, k0=0, k1=1):
    cam = synth.camera(cfg.camera)
    world = synth.BandWorld(cam.cols, cam.rows, seed, max_frames=8)
    ora = pipeline.StereoFramePointGeneratorOracle(cfg, cam, "a")
    ora.initialize(*world.pair(k0), True)
    ora.compute()
    prev = previous_points(ora, ora.framepoints())
    ora.initialize(*world.pair(k1), False)
    return cam, world, ora, prev


def _motion(cam, frames=1, noise=0.0, seed=0):
    """previous -> current for a camera x-translation of frames * B/4 (+ a small perturbation of the prior)"""
    T = np.hstack([np.eye(3), np.zeros((3, 1))])
    T[0, 3] = -frames * (-cam.bx / cam.fx) / 4
    if noise:
        rng = np.random.default_rng(seed)
        T[:, 3] += rng.normal(0, noise, 3)
        T[:, :3] = synth._rot(*rng.normal(0, noise * 0.02, 3))
    return T


@pytest.mark.parametrize("cfg_name,by_appearance,D,noise", [
    ("kitti_fast", True, 50, 0.0), ("kitti_fast", False, 15, 0.02), ("euroc", True, 50, 0.01),
    ("euroc", False, 25, 0.0), ("kitti", False, 15, 0.05)])
def test_track_matches_python_restatement(cfg_name, by_appearance, D, noise):
    cfg = configs.BY_NAME[cfg_name]
    cam, world, ora, prev = _two_frames(cfg, seed=11)
    T = _motion(cam, 1, noise, seed=5)
    # duplicates and near-duplicates of previous points compete for the same features (consumption order matters)
    extra = prev[::7].copy()
    extra["cam"][:, 0] += 0.01
    prev = np.concatenate([prev, extra, prev[::11]])
    fl, fr = ora.features_left.copy(), ora.features_right.copy()
    want_tracks, want_lost, wml, wmr = py_track(fl, fr, cam.rows, cam.cols, cam, prev, T, by_appearance, D, 38.4,
                                                ora.max_distance, cfg.minimum_disparity_pixels)
    r = ora.track(prev, T, by_appearance, D, 38.4)
    t = r["tracks"]
    assert len(t) > 0.4 * len(prev) / 1.25
    got = list(zip(t["index_previous"].tolist(), t["index_left"].tolist(), t["index_right"].tolist(),
                   t["distance"].tolist(), t["epipolar_offset"].tolist(), t["projection_left"][:, 0].astype(int).tolist(),
                   t["projection_left"][:, 1].astype(int).tolist(),
                   t["projection_right_corrected"][:, 0].astype(int).tolist(),
                   t["projection_right_corrected"][:, 1].astype(int).tolist()))
    assert got == want_tracks
    assert r["lost"].tolist() == want_lost
    assert np.array_equal(r["matched_left"], wml) and np.array_equal(r["matched_right"], wmr)
    assert r["accumulated_distance"] == float(t["distance"].sum())
    assert r["tracked_landmarks"] == len(t)
    # every feature is consumed at most once, the triangulated point uses the two matched keypoints
    assert len(set(t["index_left"].tolist())) == len(t) and len(set(t["index_right"].tolist())) == len(t)
    for q in t[:50]:
        assert np.array_equal(q["cam"], tier_a.triangulate(ora.stereo_camera, q["xl"], q["yl"], q["xr"], q["yr"]))
    # the pools compute() scans afterwards lost exactly the matched features
    assert len(ora.features_left) == len(fl) - wml.sum() and len(ora.features_right) == len(fr) - wmr.sum()


def test_track_then_compute_fills_only_free_bins():
    cfg = configs.KITTI_FAST
    cam, world, ora, prev = _two_frames(cfg, seed=4)
    r = ora.track(prev, _motion(cam), False, 15, 25.6)
    tracked = ora.tracked_points(r["tracks"])
    ora.compute(tracked)
    new = ora.framepoints()
    used_l, used_r = set(r["tracks"]["index_left"].tolist()), set(r["tracks"]["index_right"].tolist())
    assert not used_l & set(new["index_left"].tolist()) and not used_r & set(new["index_right"].tolist())
    bs = cfg.bin_size_pixels
    bins_tracked = {(int(np.rint(t["row"] / bs)), int(np.rint(t["col"] / bs))) for t in tracked}
    bins_new = {(int(np.rint(int(q["yl"]) / bs)), int(np.rint(int(q["xl"]) / bs))) for q in new}
    assert not bins_tracked & bins_new          # :383 a tracked point (previous() set) is never replaced
    assert len(new) > 50


def test_track_skips_and_degenerate_inputs():
    cfg = configs.KITTI_FAST
    cam, world, ora, prev = _two_frames(cfg, seed=2)
    p = prev[:6].copy()
    p["cam"][0] = [0.0, 0.0, 0.0]              # 0/0 -> NaN -> integer indefinite -> skipped, not lost
    p["cam"][1] = [1e9, 0.0, 1.0]              # projects far outside -> skipped
    p["cam"][2] = [0.0, 0.0, -5.0]             # behind the camera but projecting to (cx, cy): searched like any other
    p["desc_left"][3] = ~p["desc_left"][3]     # no appearance match -> lost
    T = _motion(cam)
    T[0, 3] = 0.0
    r = ora.track(p, T, True, 50, 38.4)
    assert 0 not in r["lost"] and 1 not in r["lost"] and 0 not in r["tracks"]["index_previous"]
    assert 3 in r["lost"]
    e = ora.track(prev[:0], T, True, 50, 38.4)
    assert len(e["tracks"]) == 0 and len(e["lost"]) == 0 and e["accumulated_distance"] == 0.0


def test_recover_descriptor_equals_orb_on_region_of_interest():
    """stereo_framepoint_generator.cpp:769-795: cv::ORB::compute on image(Rect(projection - 35, 71 x 71)) with the
    keypoint at (35, 35) -- the oracle takes the descriptor from the blurred frame at the projection instead"""
    cv2 = pytest.importorskip("cv2")
    cv2.setNumThreads(0)
    cfg = configs.KITTI
    cam, world, ora, prev = _two_frames(cfg, seed=9)
    lost = prev[:200].copy()
    W = _motion(cam)
    rec = ora.recover_points(lost, W, 64.0)
    assert len(rec) > 100
    orb = cv2.ORB_create()
    left, right = ora.images
    for q in rec[:60]:
        for img, x, y, want in ((left, q["xl"], q["yl"], q["desc_left"]), (right, q["xr"], q["yr"], q["desc_right"])):
            cx, cy = int(x) - 35, int(y) - 35
            roi = img[cy:cy + 71, cx:cx + 71]          # a view into the frame, like cv::Mat::operator()(Rect)
            kp = cv2.KeyPoint(35.0, 35.0, 7.0, -1.0, 0.0, 0, -1)
            kps, desc = orb.compute(roi, [kp])
            assert len(kps) == 1 and np.array_equal(desc[0], want)


def test_recover_gates():
    cfg = configs.KITTI
    cam, world, ora, prev = _two_frames(cfg, seed=9)
    W = _motion(cam)
    lost = prev[:300].copy()
    lost["has_landmark"][::3] = 0
    rec = ora.recover_points(lost, W, 38.4)
    assert len(rec) and not np.any(rec["index_lost"] % 3 == 0)          # :704-706 only landmarks
    for q in rec:
        pp = lost[q["index_lost"]]
        assert tier_a.hamming256(pp["desc_left"], q["desc_left"]) <= 38.4
        assert tier_a.hamming256(pp["desc_right"], q["desc_right"]) <= 38.4
        assert q["distance"] == tier_a.hamming256(q["desc_left"], q["desc_right"]) <= ora.max_distance
        assert q["xl"] - q["xr"] >= cfg.minimum_disparity_pixels
        assert 36 <= q["xl"] <= cam.cols - 36 and 36 <= q["yl"] <= cam.rows - 36
        assert np.array_equal(q["cam"], tier_a.triangulate(ora.stereo_camera, q["xl"], q["yl"], q["xr"], q["yr"]))
    # depth gate (:731-736) and the tracking-distance gate
    assert len(ora.recover_points(lost, W, 38.4, min_depth=0.1, max_depth=1.0)) == 0
    assert len(ora.recover_points(lost, W, -1.0)) == 0
