"""BASELINE.json configs[1]: configuration_euroc.yaml, synthetic 752x480 stereo sequence on one GPU -- framepoint
generation (2x2 detector grid, threshold feedback) + StereoUVAligner pose refinement between consecutive frames.
Per frame, in the order of PoseTracker3D::compute (reference src/position_tracking/pose_tracker_3d.cpp:80, 239, 124-126,
210): initialize -> track against all points of the previous frame -> StereoUVAligner::initialize / converge over the
tracks -> compute (new points, tracks pre-loaded from device memory).  Every stage must agree with the CPU oracle and
the aligner must recover the known camera motion."""
import numpy as np
import pytest

from oracle import pipeline, tier_a
from vslam_b200 import api, configs, synth

pytestmark = pytest.mark.gpu


def test_euroc_sequence_generation_and_alignment():
    cfg, acfg = configs.EUROC, configs.EUROC_ALIGNER
    cam = synth.camera(cfg.camera)
    world = synth.BandWorld(cam.cols, cam.rows, 3, max_frames=8)
    gen = api.StereoFramePointGenerator(cfg, cam)
    ora = pipeline.StereoFramePointGeneratorOracle(cfg, cam, "a")
    aligner = api.StereoUVAligner(acfg, max_points=4096)
    baseline_m = -cam.bx / cam.fx
    prev_g = prev_o = None
    T0 = np.hstack([np.eye(3), np.zeros((3, 1))])
    for k in range(5):
        left, right = world.pair(k)
        gen.initialize(left, right, k == 0)
        ora.initialize(left, right, k == 0)
        assert np.array_equal(gen.thresholds, ora.thresholds)
        _, dl = gen.features(0)
        _, dr = gen.features(1)
        parts_g, parts_o = [], []
        if prev_g is not None:
            # constant-velocity guess = identity (pose_tracker_3d.cpp:41-47 before the first estimate): widest window
            got = gen.track(prev_g, T0, True, 50, 38.4)
            want = ora.track(prev_o, T0, True, 50, 38.4)
            tr = got["tracks"]
            assert len(tr) == len(want["tracks"]) > 150
            assert np.array_equal(tr["index_left"], want["tracks"]["index_left"])
            assert np.array_equal(tr["camera"], want["tracks"]["cam"]) and np.array_equal(got["lost"], want["lost"])
            # StereoUVAligner::initialize (stereouv_aligner.cpp:10-69) over frame->points() == the tracks
            moving = np.ascontiguousarray(prev_g["camera_left"][tr["index_previous"]])
            fixed = np.stack([tr["xl"], tr["yl"], tr["xr"], tr["yr"]], 1).astype(np.float64)
            omega = np.ones(len(tr))
            wt = np.minimum(acfg.maximum_reliable_depth_meters / tr["camera"][:, 2], 1.0)       # :59-63
            aligner.initialize(moving, fixed, omega, wt, cam.K, cam.baseline, cam.rows, cam.cols, T0)
            aligner.converge()
            cpu = tier_a.Aligner("stereouv", moving, fixed, omega, wt, cam.K, cam.baseline, cam.rows, cam.cols,
                                 acfg.minimum_reliable_depth_meters, acfg.maximum_error_kernel)
            r = cpu.converge(T0, acfg.damping, acfg.error_delta_for_convergence, acfg.maximum_number_of_iterations,
                             acfg.minimum_number_of_inliers)
            T = aligner.previousToCurrent()
            assert aligner.has_system_converged == r["converged"] and aligner.number_of_rounds == r["rounds"]
            dR = T[:, :3] @ r["T"][:, :3].T
            assert np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1)) <= 1e-6          # parity with the oracle
            assert np.linalg.norm(T[:, 3] - r["T"][:, 3]) <= 1e-5
            # the camera moved by a quarter baseline along +x: previous -> current is a translation by -B/4
            assert np.linalg.norm(T[:, 3] - [-baseline_m / 4, 0, 0]) < 2e-3
            assert np.abs(T[:, :3] - np.eye(3)).max() < 2e-3
            assert aligner.numberOfInliers() > 0.8 * len(tr)
            ora.compute(ora.tracked_points(want["tracks"]))
            cur = gen.compute(api.TRACKED_FROM_LAST_TRACK)
            parts_g, parts_o = [tr], [want["tracks"]]
        else:
            ora.compute()
            cur = gen.compute()
        want_new = ora.framepoints()
        assert len(cur) == len(want_new) and np.array_equal(cur["index_left"], want_new["index_left"])
        assert np.array_equal(cur["camera"], want_new["cam"])
        parts_g.append(cur)
        parts_o.append(want_new)
        prev_g = api.make_previous_points(parts_g, dl, dr)
        prev_o = prev_g.view(tier_a.PREVIOUS_POINT)
    gen.close()
    aligner.close()
