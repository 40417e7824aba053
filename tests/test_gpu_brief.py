"""The BRIEF-32 descriptor branch (reference base_framepoint_generator.cpp:186, cv::xfeatures2d::BriefDescriptorExtractor
create(32)) on the device, with a supplied test table, against the CPU oracle: bit-exact keypoint sets (28 px border
instead of ORB's 31), descriptors, matches, framepoints, tracks and recovered points.  "Parity unpinned" against
opencv_contrib itself (not in this image): see tests/test_oracle_brief.py."""
import dataclasses

import numpy as np
import pytest

from oracle import pipeline
from vslam_b200 import api, configs, synth

from test_gpu_track import _as_api, _same_points, _same_tracks
from test_oracle_track import _motion, previous_points

pytestmark = pytest.mark.gpu


def _cfg(name):
    return dataclasses.replace(configs.BY_NAME[name], descriptor_type="BRIEF", brief_tests=synth.brief_test_table(32))


@pytest.mark.parametrize("name", ["kitti", "euroc"])
def test_brief32_pipeline_matches_oracle(name):
    cfg = _cfg(name)
    cam = synth.camera(cfg.camera)
    world = synth.BandWorld(cam.cols, cam.rows, 5, max_frames=4)
    gen = api.StereoFramePointGenerator(cfg, cam)
    ora = pipeline.StereoFramePointGeneratorOracle(cfg, cam, "a")
    left, right = world.pair(0)
    nl, nr = gen.initialize(left, right, True)
    ora.initialize(left, right, True)
    assert (nl, nr) == (len(ora.kps_left), len(ora.kps_right))
    for side, (kps, desc) in enumerate(((ora.kps_left, ora.desc_left), (ora.kps_right, ora.desc_right))):
        k, d = gen.features(side)
        assert np.array_equal(k["x"], kps["x"]) and np.array_equal(k["y"], kps["y"]) and np.array_equal(d, desc)
        assert k["x"].min() >= 28 and (k["x"] < 31).any() or name == "euroc"      # the 28 px border admits more keypoints
    ora.compute()
    fps = gen.compute()
    _same_points(fps, ora.framepoints())
    assert len(fps) > 300
    # next frame: track, compute, recoverPoints
    prev = previous_points(ora, ora.framepoints())
    left, right = world.pair(1)
    gen.initialize(left, right, False)
    ora.initialize(left, right, False)
    T = _motion(cam)
    want = ora.track(prev, T, False, 15, 38.4)
    got = gen.track(_as_api(prev), T, False, 15, 38.4)
    _same_tracks(got, want)
    assert len(want["tracks"]) > 200
    ora.compute(ora.tracked_points(want["tracks"]))
    _same_points(gen.compute(api.TRACKED_FROM_LAST_TRACK), ora.framepoints())
    lost = prev.copy()
    lost["keypoint_size"][::7] = 5.5          # border 27.5 < 28: dropped by the extractor's border filter
    rec_w = ora.recover_points(lost, T, 64.0)
    rec_g = gen.recover_points(_as_api(lost), T, 64.0)
    assert len(rec_g) == len(rec_w) > 100 and not np.any(rec_w["index_lost"] % 7 == 0)
    for f, g in (("index_lost", "index_lost"), ("distance", "distance"), ("xl", "xl"), ("xr", "xr"), ("camera", "cam"),
                 ("descriptor_left", "desc_left"), ("descriptor_right", "desc_right")):
        assert np.array_equal(rec_g[f], rec_w[g]), f
    gen.close()


def test_brief32_batched_matches_single():
    cfg = _cfg("kitti_fast")
    cam = synth.camera(cfg.camera)
    left, right = synth.band_world_batch(cfg.camera, range(3))
    gen = api.StereoFramePointGenerator(cfg, cam, max_batch=3)
    out, counts = gen.batch_process(left, right, True)
    for i in range(3):
        o = pipeline.StereoFramePointGeneratorOracle(cfg, cam, "a").initialize(left[i], right[i], True)
        o.compute()
        _same_points(out[i, :counts[i]], o.framepoints())
    gen.close()


def test_brief32_configuration_errors():
    cam = synth.camera("kitti")
    with pytest.raises(api.VslamError):       # offsets beyond the 48 px patch
        bad = synth.brief_test_table(1).copy()
        bad[3, 2] = 25
        api.StereoFramePointGenerator(dataclasses.replace(configs.KITTI, descriptor_type="BRIEF", brief_tests=bad), cam)
    c = api.make_config(configs.KITTI, cam)
    c.descriptor_type = 1                     # BRIEF without a table
    import ctypes as C
    h = C.c_void_p()
    assert api.lib().vslam_fpg_create(C.byref(c), 0, C.byref(h)) == -1
    c.descriptor_type = 7
    assert api.lib().vslam_fpg_create(C.byref(c), 0, C.byref(h)) == -1
