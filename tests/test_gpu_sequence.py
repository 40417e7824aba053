"""BASELINE.json configs[1]: configuration_euroc.yaml, synthetic 752x480 stereo sequence on one GPU -- framepoint
generation (2x2 detector grid, threshold feedback) + StereoUVAligner pose refinement between consecutive frames.
The frame-to-frame correspondences come from the band world's known geometry (track() itself is a SURVEY 8(f)
"next" row); the aligner must agree with the CPU oracle and recover the known camera motion."""
import numpy as np
import pytest

from oracle import pipeline, tier_a
from vslam_b200 import api, configs, synth

pytestmark = pytest.mark.gpu


def _correspond(world, prev, cur):
    """framepoints of `cur` whose left pixel is where the band geometry puts a framepoint of `prev`"""
    lut = {(int(p["xl"]), int(p["yl"])): i for i, p in enumerate(prev)}
    pairs = []
    for j, q in enumerate(cur):
        band = min(int(q["yl"]) // synth.BAND_ROWS, len(world.band_disparity) - 1)
        shift = world.band_disparity[band] // 4
        i = lut.get((int(q["xl"]) + shift, int(q["yl"])))
        if i is not None:
            pairs.append((i, j))
    return np.array(pairs)


def test_euroc_sequence_generation_and_alignment():
    cfg, acfg = configs.EUROC, configs.EUROC_ALIGNER
    cam = synth.camera(cfg.camera)
    world = synth.BandWorld(cam.cols, cam.rows, 3, max_frames=8)
    gen = api.StereoFramePointGenerator(cfg, cam)
    ora = pipeline.StereoFramePointGeneratorOracle(cfg, cam, "a")
    aligner = api.StereoUVAligner(acfg, max_points=4096)
    baseline_m = -cam.bx / cam.fx
    prev = None
    for k in range(5):
        left, right = world.pair(k)
        gen.initialize(left, right, k == 0)
        ora.initialize(left, right, k == 0)
        ora.compute()
        cur = gen.compute()
        want = ora.framepoints()
        assert len(cur) == len(want) and np.array_equal(cur["index_left"], want["index_left"])
        assert np.array_equal(cur["camera"], want["cam"]) and np.array_equal(gen.thresholds, ora.thresholds)
        if prev is not None:
            pairs = _correspond(world, prev, cur)
            assert len(pairs) > 150
            moving = np.ascontiguousarray(prev["camera"][pairs[:, 0]])
            c = cur[pairs[:, 1]]
            fixed = np.stack([c["xl"], c["yl"], c["xr"], c["yr"]], 1).astype(np.float64)
            omega = np.ones(len(pairs))
            wt = np.minimum(acfg.maximum_reliable_depth_meters / c["camera"][:, 2], 1.0)   # stereouv_aligner.cpp:59-63
            T0 = np.hstack([np.eye(3), np.zeros((3, 1))])
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
            assert aligner.numberOfInliers() > 0.8 * len(pairs)
        prev = cur
    gen.close()
    aligner.close()
