"""vslam_fpg_frame_step -- one tracked frame as ONE device pass (graph launch + one synchronisation) -- against the
stepwise calls in the tracker's order (initialize -> track -> StereoUVAligner initialize + converge -> _prunePoints ->
compute -> points() of the frame for the next track()), which tests/test_gpu_track.py, test_gpu_aligner.py and
test_gpu_fpg.py pin to the oracle stage by stage: every output of every frame must be bit-identical -- tracks, lost
list, aligner errors / inliers / pose / round count, the prune decision, the new framepoints and the 128-byte points()
records with their descriptors."""
import dataclasses

import numpy as np
import pytest

from vslam_b200 import api, configs, synth

pytestmark = pytest.mark.gpu

TRACK_FIELDS = ("index_previous", "xl", "yl", "xr", "yr", "distance", "epipolar_offset", "projection_left",
                "projection_right", "projection_right_corrected", "camera")
POINT_FIELDS = ("xl", "yl", "xr", "yr", "distance", "epipolar_offset", "camera")


def _prior(cam, error=0.0):
    T = np.hstack([np.eye(3), np.zeros((3, 1))])
    T[0, 3] = -(-cam.bx / cam.fx) / 4 + error      # the band world moves a quarter baseline per frame along x
    return T


class Stepwise:
    """the tracker's per-frame order through the stage-by-stage C ABI (what tools/sequence_runner.cpp does)"""

    def __init__(self, cfg, acfg, cam, D, max_distance, min_track_length=1):
        self.cfg, self.acfg, self.cam, self.D, self.max_distance = cfg, acfg, cam, D, max_distance
        self.min_track_length = min_track_length
        self.gen = api.StereoFramePointGenerator(cfg, cam)
        self.aligner = api.StereoUVAligner(acfg, max_points=8192)
        self.previous = np.zeros(0, api.PREVIOUS_POINT)
        self.estimates = None          # LANDMARK_ESTIMATE per previous point (stereouv_aligner.cpp:43-51) or None

    def step(self, left, right, localizing, T_prior):
        gen, cam, acfg = self.gen, self.cam, self.acfg
        out = {}
        out["n_left"], out["n_right"] = gen.initialize(left, right, localizing)
        prev = self.previous
        tr = gen.track(prev, T_prior, False, self.D, self.max_distance)
        tracks = tr["tracks"]
        out.update(n_previous=len(prev), n_tracked=len(tracks), lost=tr["lost"], n_tracked_landmarks=tr["tracked_landmarks"],
                   average_descriptor_distance=tr["average_descriptor_distance"])
        T = np.array(T_prior, np.float64)
        keep = np.zeros(0, bool)
        out.update(aligner_rounds=0, aligner_converged=0, errors=np.zeros(0), inliers=np.zeros(0, bool))
        if len(tracks):
            moving = np.ascontiguousarray(prev["camera_left"][tracks["index_previous"]])
            omega = np.ones(len(tracks))
            if self.estimates is not None:     # a landmark estimate is preferred, its information scaled (:43-51)
                est = self.estimates[tracks["index_previous"]]
                has = est["information_scale"] != 0
                moving[has] = est["camera"][has]
                omega[has] = est["information_scale"][has]
            fixed = np.stack([tracks["xl"], tracks["yl"], tracks["xr"], tracks["yr"]], 1).astype(np.float64)
            q = acfg.maximum_reliable_depth_meters / tracks["camera"][:, 2]
            wt = np.where(1.0 < q, 1.0, q) if acfg.enable_inverse_depth_as_information else np.ones(len(tracks))
            self.aligner.initialize(moving, fixed, omega, wt, cam.K, cam.baseline, cam.rows, cam.cols, T)
            self.aligner.converge(fused=True)
            T = np.array(self.aligner.previousToCurrent(), np.float64)
            keep = gen.prune_tracks(self.aligner, acfg.maximum_error_kernel)
            out.update(aligner_rounds=self.aligner.number_of_rounds,
                       aligner_converged=int(self.aligner.has_system_converged),
                       errors=self.aligner.errors(), inliers=self.aligner.inliers())
        out["previous_to_current"] = np.asarray(T).reshape(3, 4)
        out["kept"] = keep
        kept_tracks = tracks[keep] if len(tracks) else tracks
        out["tracks"] = kept_tracks
        points = gen.compute(api.TRACKED_FROM_LAST_TRACK)
        out["points"] = points
        out["n_matches"] = gen.number_of_matches
        _, dl = gen.features(0)
        _, dr = gen.features(1)
        nxt = api.make_previous_points([kept_tracks, points], dl, dr)
        length = np.concatenate([prev["reserved"][kept_tracks["index_previous"]] + 1, np.ones(len(points), np.int32)])
        nxt["reserved"] = length
        nxt["has_landmark"] = length >= self.min_track_length
        out["frame_points"] = nxt
        self.previous = nxt
        self.estimates = None          # estimates belong to the points they were given for
        return out

    def close(self):
        self.gen.close()
        self.aligner.close()


def _compare(k, got, want, gen_fused, single_region):
    ctx = "frame %d" % k
    for f in ("n_left", "n_right", "n_previous", "n_tracked", "n_tracked_landmarks", "n_matches", "aligner_rounds",
              "aligner_converged"):
        assert got[f] == want[f], (ctx, f, got[f], want[f])
    assert np.array_equal(got["lost"], want["lost"]), ctx
    if want["n_tracked"]:
        assert got["average_descriptor_distance"] == want["average_descriptor_distance"], ctx
    else:
        assert np.isnan(got["average_descriptor_distance"]), ctx
    assert np.array_equal(got["errors"], want["errors"]), ctx
    assert np.array_equal(got["inliers"], want["inliers"]), ctx
    assert np.array_equal(got["kept"], want["kept"]), ctx
    assert np.array_equal(got["previous_to_current"], want["previous_to_current"]), ctx      # bit-identical pose
    assert got["n_tracks"] == len(want["tracks"]) and got["n_new_points"] == len(want["points"]), ctx
    for f in TRACK_FIELDS:
        assert np.array_equal(got["tracks"][f], want["tracks"][f]), (ctx, f)
    for f in POINT_FIELDS:
        assert np.array_equal(got["points"][f], want["points"][f]), (ctx, f)
    if single_region:   # the device's (row, col) order IS the reference's keypoint order
        for f in ("index_left", "index_right"):
            assert np.array_equal(got["tracks"][f], want["tracks"][f]), (ctx, f)
            assert np.array_equal(got["points"][f], want["points"][f]), (ctx, f)
    a, b = got["frame_points"], want["frame_points"]
    assert len(a) == len(b), ctx
    for f in ("camera_left", "world", "descriptor_left", "descriptor_right", "epipolar_offset", "has_landmark",
              "keypoint_size", "reserved"):
        assert np.array_equal(a[f], b[f]), (ctx, f)


@pytest.mark.parametrize("cfg_name,frames,prior_error,min_track_length", [
    ("kitti", 8, 0.0, 1), ("kitti", 5, 0.05, 3), ("euroc", 6, 0.0, 1), ("kitti_fast", 5, 0.02, 2), ("hd", 4, 0.0, 1)])
def test_frame_step_equals_the_stepwise_calls(cfg_name, frames, prior_error, min_track_length):
    cfg, acfg = configs.BY_NAME[cfg_name], configs.ALIGNER_BY_NAME[cfg_name]
    cam = synth.camera(cfg.camera)
    world = synth.BandWorld(cam.cols, cam.rows, 31 + frames, max_frames=frames)
    D, max_distance = 25, 40.0
    ref = Stepwise(cfg, acfg, cam, D, max_distance, min_track_length)
    gen = api.StereoFramePointGenerator(cfg, cam)
    assert gen.frame_step_capacity() >= 2048
    gen.frame_step_reset()
    T = _prior(cam, prior_error)
    total_tracks = total_pruned = 0
    for k in range(frames):
        left, right = world.pair(k)
        want = ref.step(left, right, k == 0, T)
        got = gen.frame_step(left, right, k == 0, T, acfg, False, D, max_distance, min_track_length)
        _compare(k, got, want, gen, gen.number_of_detectors == 1)
        assert np.array_equal(gen.thresholds, ref.gen.thresholds), k
        total_tracks += got["n_tracks"]
        total_pruned += got["n_tracked"] - got["n_tracks"]
    assert total_tracks > 100 * (frames - 1)
    if prior_error:
        assert total_pruned > 0            # the prune rule was exercised
    assert gen.graph_launch_count >= frames - 1
    ref.close()
    gen.close()


def test_frame_step_with_previous_points_from_the_host_and_reset():
    cfg, acfg = configs.KITTI, configs.KITTI_ALIGNER
    cam = synth.camera(cfg.camera)
    world = synth.BandWorld(cam.cols, cam.rows, 77, max_frames=3)
    ref = Stepwise(cfg, acfg, cam, 25, 40.0)
    T = _prior(cam)
    ref.step(*world.pair(0), True, T)
    gen = api.StereoFramePointGenerator(cfg, cam)
    # the generator's thresholds follow the same one-frame history; the previous points come from the host
    gen.initialize(*world.pair(0), True)
    gen.compute()
    gen.frame_step_set_previous(ref.previous)
    want = ref.step(*world.pair(1), False, T)
    got = gen.frame_step(*world.pair(1), False, T, acfg, False, 25, 40.0)
    _compare(1, got, want, gen, True)
    assert got["n_tracks"] > 200
    # a new sequence: nothing to track against, the frame is a first frame again
    gen.frame_step_reset()
    got = gen.frame_step(*world.pair(2), True, T, acfg, False, 25, 40.0)
    assert got["n_previous"] == 0 and got["n_tracked"] == 0 and got["n_tracks"] == 0 and got["aligner_rounds"] == 0
    assert np.array_equal(got["previous_to_current"], T)
    assert got["n_new_points"] > 200 and len(got["frame_points"]) == got["n_new_points"]
    ref.close()
    gen.close()


def test_frame_step_needs_binning():
    cfg = dataclasses.replace(configs.KITTI, enable_keypoint_binning=False)
    cam = synth.camera(cfg.camera)
    gen = api.StereoFramePointGenerator(cfg, cam)
    assert gen.frame_step_capacity() == 0
    left, right = synth.band_world_pair(cfg.camera, 3)
    with pytest.raises(api.VslamError):
        gen.frame_step(left, right, True, _prior(cam), configs.KITTI_ALIGNER, False, 25, 40.0)
    gen.close()


def test_prefetched_frames_give_the_same_results_and_the_inbox_rules_hold():
    """vslam_fpg_frame_step_prefetch: the images of frame k + 1 travel while frame k runs (two inbox buffers, a copy
    stream); the results equal the stepwise calls frame by frame, mixed with frames that bring their own images"""
    cfg, acfg = configs.KITTI, configs.KITTI_ALIGNER
    cam = synth.camera(cfg.camera)
    frames = 7
    world = synth.BandWorld(cam.cols, cam.rows, 91, max_frames=frames)
    pairs = [world.pair(k) for k in range(frames)]
    D, max_distance = 25, 40.0
    ref = Stepwise(cfg, acfg, cam, D, max_distance)
    gen = api.StereoFramePointGenerator(cfg, cam)
    gen.frame_step_reset()
    T = _prior(cam)
    with pytest.raises(api.VslamError):          # nothing staged
        gen.frame_step(None, None, True, T, acfg, False, D, max_distance)
    own_images = 4                               # this frame is not prefetched: it brings its own images
    gen.frame_step_prefetch(*pairs[0])
    for k in range(frames):
        ahead = k + 1 < frames and k + 1 != own_images and k != own_images
        if ahead:                                # frame k + 1 travels while frame k runs
            gen.frame_step_prefetch(*pairs[k + 1])
        want = ref.step(*pairs[k], k == 0, T)
        if k == own_images:
            got = gen.frame_step(*pairs[k], False, T, acfg, False, D, max_distance)
            gen.frame_step_prefetch(*pairs[k + 1])
        else:
            got = gen.frame_step(None, None, k == 0, T, acfg, False, D, max_distance)
        _compare(k, got, want, gen, True)
        assert np.array_equal(gen.thresholds, ref.gen.thresholds), k
    # two staged pairs at most; a frame with images while a pair is staged is refused; reset drops staged pairs
    gen.frame_step_prefetch(*pairs[0])
    gen.frame_step_prefetch(*pairs[1])
    with pytest.raises(api.VslamError):
        gen.frame_step_prefetch(*pairs[2])
    with pytest.raises(api.VslamError):
        gen.frame_step(*pairs[2], False, T, acfg, False, D, max_distance)
    gen.frame_step_reset()
    with pytest.raises(api.VslamError):
        gen.frame_step(None, None, True, T, acfg, False, D, max_distance)
    ref.close()
    gen.close()


@pytest.mark.parametrize("switch", ["VSLAM_NO_FRAME_BRANCHES", "VSLAM_FRAME_STEP_SYNC"])
def test_frame_step_variants_one_chain_and_stream_synchronize(switch, monkeypatch):
    """the A/B switches of the fused frame -- one chain of kernels instead of parallel branches (which also runs the bin
    selection as ONE kernel), cudaStreamSynchronize instead of the polled completion word -- give the same frames"""
    monkeypatch.setenv(switch, "1")     # read when the handle (the fused frame's state) is created
    cfg, acfg = configs.KITTI, configs.KITTI_ALIGNER
    cam = synth.camera(cfg.camera)
    frames = 5
    world = synth.BandWorld(cam.cols, cam.rows, 57, max_frames=frames)
    D, max_distance = 25, 40.0
    ref = Stepwise(cfg, acfg, cam, D, max_distance)
    gen = api.StereoFramePointGenerator(cfg, cam)
    gen.frame_step_reset()
    T = _prior(cam, 0.03)
    for k in range(frames):
        left, right = world.pair(k)
        want = ref.step(left, right, k == 0, T)
        got = gen.frame_step(left, right, k == 0, T, acfg, False, D, max_distance)
        _compare(k, got, want, gen, True)
    ref.close()
    gen.close()


def test_frame_step_when_every_previous_point_is_lost():
    """a motion prior that is far off: previous points exist, none is tracked -- the aligner has nothing to align (the
    control block still carries the prior, zero rounds), nothing survives, the frame consists of new points only, and the
    sequence carries on from there"""
    cfg, acfg = configs.KITTI, configs.KITTI_ALIGNER
    cam = synth.camera(cfg.camera)
    world = synth.BandWorld(cam.cols, cam.rows, 63, max_frames=4)
    D, max_distance = 25, 40.0
    ref = Stepwise(cfg, acfg, cam, D, max_distance)
    gen = api.StereoFramePointGenerator(cfg, cam)
    gen.frame_step_reset()
    good = _prior(cam)
    off = good.copy()
    off[1, 3] = 40.0                    # 40 m sideways: every projection leaves the image or finds nothing nearby
    for k, T in enumerate([good, off, good, good]):
        want = ref.step(*world.pair(k), k == 0, T)
        got = gen.frame_step(*world.pair(k), k == 0, T, acfg, False, D, max_distance)
        _compare(k, got, want, gen, True)
        if k == 1:
            assert got["n_previous"] > 300 and got["n_tracked"] == 0 and got["n_tracks"] == 0
            assert got["aligner_rounds"] == 0 and np.array_equal(got["previous_to_current"], off)
            assert got["n_new_points"] > 300
        if k >= 2:
            assert got["n_tracks"] > 200
    ref.close()
    gen.close()


ERR_CAPACITY = -3   # VSLAM_ERR_CAPACITY (include/vslam_b200.h)


def test_frame_step_capacity():
    """more points than the fused path holds: VSLAM_ERR_CAPACITY (the stepwise calls have no such limit), the device-
    resident points are dropped, and the next frame starts a new sequence"""
    cfg = dataclasses.replace(configs.HD, name="hd_dense", bin_size_pixels=6, detector_threshold_minimum=5,
                              detector_threshold_maximum=20)
    acfg = configs.KITTI_ALIGNER
    cam = synth.camera(cfg.camera)
    world = synth.BandWorld(cam.cols, cam.rows, 17, max_frames=3)
    gen = api.StereoFramePointGenerator(cfg, cam, max_keypoints=32768)
    cap = gen.frame_step_capacity()
    T = _prior(cam)
    # from the host: more previous points than the capacity
    with pytest.raises(api.VslamError) as e:
        gen.frame_step_set_previous(np.zeros(cap + 1, api.PREVIOUS_POINT))
    assert e.value.code == ERR_CAPACITY
    # from a frame: 6-pixel bins on a 1920 x 1080 pair hold far more new points than the capacity
    gen.frame_step_reset()
    gen.thresholds = np.full(gen.number_of_detectors, 5.0)
    with pytest.raises(api.VslamError) as e:
        gen.frame_step(*world.pair(0), True, T, acfg, False, 25, 40.0)
    assert e.value.code == ERR_CAPACITY
    # the same pair through the stepwise calls is fine, and shows that the limit was the reason
    gen.initialize(*world.pair(0), True)
    assert len(gen.compute()) > cap
    gen.close()
    # a handle with room: the frame after a refused one is a first frame again
    gen = api.StereoFramePointGenerator(configs.HD, cam)
    gen.frame_step_reset()
    got = gen.frame_step(*world.pair(1), True, T, acfg, False, 25, 40.0)
    assert got["n_previous"] == 0 and 0 < got["n_new_points"] <= cap
    gen.close()


def test_frame_step_is_deterministic_under_repetition():
    """compute-sanitizer is closed on this GPU pool; a race in the cluster kernels, the shared-memory claims of the track
    resolver or the parallel branches of the frame graph would show up as run-to-run differences: the same four frames,
    40 times over (fresh sequence each time), must give the same bytes"""
    import hashlib
    cfg, acfg = configs.KITTI, configs.KITTI_ALIGNER
    cam = synth.camera(cfg.camera)
    world = synth.BandWorld(cam.cols, cam.rows, 29, max_frames=4)
    pairs = [world.pair(k) for k in range(4)]
    T = _prior(cam, 0.02)
    gen = api.StereoFramePointGenerator(cfg, cam)
    first = None
    for rep in range(40):
        gen.frame_step_reset()
        gen.thresholds = np.full(gen.number_of_detectors, 50.0)
        h = hashlib.sha256()
        for k, (left, right) in enumerate(pairs):
            got = gen.frame_step(left, right, k == 0, T, acfg, False, 25, 40.0)
            for key in ("tracks", "kept", "errors", "inliers", "lost", "points", "frame_points", "previous_to_current",
                        "information"):
                h.update(np.ascontiguousarray(got[key]).tobytes())
            h.update(np.array([got[f] for f in ("n_left", "n_right", "n_tracked", "n_tracks", "n_new_points", "n_matches",
                                                "aligner_rounds", "aligner_inliers")], np.int64).tobytes())
        if first is None:
            first = h.hexdigest()
            assert got["n_tracks"] > 200
        assert h.hexdigest() == first, rep
    gen.close()


def test_frame_step_with_landmark_estimates():
    """StereoUVAligner::initialize prefers a landmark estimate and scales its information with the landmark's updates
    (stereouv_aligner.cpp:43-51): the host pushes the estimates for the points() the device holds, the next fused frame
    aligns with them -- bit-identical to the stepwise calls fed the same moving points and information"""
    cfg, acfg = configs.KITTI, configs.KITTI_ALIGNER
    cam = synth.camera(cfg.camera)
    frames = 6
    world = synth.BandWorld(cam.cols, cam.rows, 83, max_frames=frames)
    D, max_distance = 25, 40.0
    ref = Stepwise(cfg, acfg, cam, D, max_distance)
    gen = api.StereoFramePointGenerator(cfg, cam)
    gen.frame_step_reset()
    T = _prior(cam, 0.02)
    rng = np.random.default_rng(4)
    with_estimates = 0
    for k in range(frames):
        want = ref.step(*world.pair(k), k == 0, T)
        got = gen.frame_step(*world.pair(k), k == 0, T, acfg, False, D, max_distance)
        _compare(k, got, want, gen, True)
        points = got["frame_points"]
        # "landmarks": points tracked for at least two frames get an estimate close to their camera coordinates and the
        # reference's information scale 1 + log(number of updates); frame 3 deliberately gives none
        est = np.zeros(len(points), api.LANDMARK_ESTIMATE)
        has = (points["reserved"] >= 2) & (rng.random(len(points)) < 0.8)
        est["camera"][has] = points["camera_left"][has] + rng.normal(0, 0.01, (int(has.sum()), 3))
        est["information_scale"][has] = 1.0 + np.log(points["reserved"][has].astype(np.float64))
        if k != 3:
            gen.frame_step_set_landmark_estimates(est)
            ref.estimates = est
            with_estimates += int(has.sum())
    assert with_estimates > 500
    # one entry per point, or the call is refused
    with pytest.raises(api.VslamError):
        gen.frame_step_set_landmark_estimates(est[:-1])
    ref.close()
    gen.close()
