"""The drop-in claim, executed: the reference's own, UNMODIFIED PoseTracker3D / WorldMap / Frame / Landmark code
(oracle/_ref, compiled from /root/reference) drives adapters/GpuStereoFramePointGenerator and GpuStereoUVAligner -- i.e.
libvslam_b200.so through include/vslam_b200.h -- exactly where it drives its own StereoFramePointGenerator and
StereoUVAligner (src/system/slam_assembly.cpp:48-76), and the two systems are run side by side
on the same synthetic sequences.

BASELINE.json configs[0] (configuration_kitti.yaml, 200-frame 1241 x 376 sequence through the tracker executables/app
runs per frame) and configs[1] (configuration_euroc.yaml sequence, generation + StereoUVAligner tracking) are the two
long cases.

Bar (BASELINE.json north_star): keypoints, descriptors, matches, tracks, lost / recovered points, landmark bookkeeping:
bit-exact.  Triangulated points: bit-exact (1e-4 relative allowed).  Pose per frame: <= 1e-6 rad, <= 1e-5 m allowed;
asserted at 1e-9 (the only difference between the arms is the summation order of H and b on the device).
"""
import os

import numpy as np
import pytest

from oracle import ref
from vslam_b200 import configs, synth

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref.gpu_available(), reason="oracle/_ref/libvslam_ref_gpu.so is not built")]

YAML = {"kitti": "configuration_kitti.yaml", "kitti_fast": "configuration_kitti_fast.yaml",
        "euroc": "configuration_euroc.yaml"}
HAVE_YAML = os.path.isdir(os.path.join(ref.REFERENCE_ROOT, "configurations"))
EXACT = ("xl", "yl", "xr", "yr", "row", "col", "epipolar_offset", "index_previous", "disparity", "distance", "cam",
         "projection_left", "projection_right", "projection_right_corrected", "has_landmark", "track_length",
         "landmark_updates", "desc_left", "desc_right")
CLOSE = ("robot", "world", "landmark_world")


def _pair_of_sessions(name, **overrides):
    """the reference with its CPU classes, and the reference with the GPU adapters in their place"""
    cfg = configs.BY_NAME[name]
    cam = synth.camera(cfg.camera)
    if HAVE_YAML:
        make = lambda gpu: ref.Session(cam, YAML[name], gpu=gpu, **overrides)
    else:   # the GPU box has no /root/reference: the values the YAML parses to (tests/test_oracle_vs_ref.py pins them)
        values = ref.effective_values(name)
        values.update(overrides)
        make = lambda gpu: ref.Session(cam, None, gpu=gpu, **values)
    return cfg, cam, make(False).configure(), make(True).configure()


def _assert_points(a, b, where):
    assert len(a) == len(b), where
    for key in EXACT:
        assert np.array_equal(a[key], b[key]), (where, key)
    for key in CLOSE:
        assert np.abs(a[key] - b[key]).max(initial=0.0) <= 1e-9 * max(1.0, np.abs(a[key]).max(initial=0.0)), (where, key)


def _pose_error(A, B):
    dR = A[:, :3] @ B[:, :3].T        # sin(angle) from the skew part: arccos of the trace cannot resolve below 1e-8
    w = np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]]) / 2
    return float(np.arcsin(min(np.linalg.norm(w), 1.0))), float(np.linalg.norm(A[:, 3] - B[:, 3]))


@pytest.mark.parametrize("name,frames,seed", [("kitti", 200, 1), ("euroc", 60, 3), ("kitti_fast", 40, 5)])
def test_pose_tracker_with_the_gpu_adapters_equals_the_cpu_reference(name, frames, seed):
    """PoseTracker3D::compute() per frame: initialize -> track (+ recursive retries) -> aligner initialize / converge ->
    prune -> recoverPoints -> landmark updates -> compute; state machine Localizing -> Tracking included"""
    cfg, cam, cpu, gpu = _pair_of_sessions(name)
    world = synth.BandWorld(cam.cols, cam.rows, seed, max_frames=frames)
    baseline_m = -cam.bx / cam.fx
    worst = [0.0, 0.0]
    for k in range(frames):
        left, right = world.pair(k)
        n_cpu, n_gpu = cpu.process(left, right), gpu.process(left, right)
        where = f"{name} frame {k}"
        assert n_cpu == n_gpu > 100, where
        assert cpu.status() == gpu.status(), where
        assert cpu.counts() == gpu.counts(), where
        _assert_points(cpu.points(), gpu.points(), where)
        for side in (0, 1):
            (ka, da), (kb, db) = cpu.features(side), gpu.features(side)
            assert np.array_equal(ka, kb) and np.array_equal(da, db), where
        rad, metres = _pose_error(cpu.pose(), gpu.pose())
        worst = [max(worst[0], rad), max(worst[1], metres)]
        assert rad <= 1e-9 and metres <= 1e-9, (where, rad, metres)
    # both arms recover the known motion of the band world: k * B / 4 along +x
    T = gpu.pose()
    assert cpu.status() == 1
    assert abs(T[0, 3] - (frames - 1) * baseline_m / 4) < 0.01 * frames * baseline_m / 4 + 2e-3
    assert abs(T[1, 3]) < 5e-3 and abs(T[2, 3]) < 5e-3 + 2e-4 * frames
    print(f"{name}: {frames} frames, worst pose difference {worst[0]:.2e} rad {worst[1]:.2e} m")
    cpu.close()
    gpu.close()


@pytest.mark.parametrize("name,by_appearance,distance", [("kitti", False, 15), ("euroc", True, 50)])
def test_generator_stage_by_stage_with_the_gpu_adapter(name, by_appearance, distance):
    """the calls PoseTracker3D makes on the generator, one by one, with landmarks in play: the object graph the adapter
    materialises (adapters/gpu_stereo_framepoint_generator.cpp) equals the one the reference's own generator builds"""
    cfg, cam, cpu, gpu = _pair_of_sessions(name)
    world = synth.BandWorld(cam.cols, cam.rows, 17, max_frames=8)
    T = np.hstack([np.eye(3), np.zeros((3, 1))])
    T[0, 3] = -(-cam.bx / cam.fx) / 4
    pose = np.hstack([np.eye(3), np.zeros((3, 1))])
    for k in range(6):
        left, right = world.pair(k)
        for s in (cpu, gpu):
            s.initialize(left, right, tracking=k >= 2)
        where = f"{name} frame {k}"
        if k:
            pose = pose.copy()
            pose[0, 3] += (-cam.bx / cam.fx) / 4
            out = []
            for s in (cpu, gpu):
                s.set_tracking(distance, 25.6 + 5 * k)
                r = s.track(T, by_appearance)
                s.set_pose(pose)
                r["recovered"] = s.recover()
                out.append(r)
            a, b = out
            assert a["n_tracks"] == b["n_tracks"] > 50 and np.array_equal(a["lost"], b["lost"]), where
            assert a["number_of_tracked_landmarks"] == b["number_of_tracked_landmarks"], where
            assert a["average_descriptor_distance"] == b["average_descriptor_distance"], where
            assert a["recovered"] == b["recovered"], where
            if k >= 3:
                assert a["recovered"] > 0, where
            _assert_points(cpu.points(), gpu.points(), where + " after track + recoverPoints")
        assert cpu.compute() == gpu.compute(), where
        _assert_points(cpu.points(), gpu.points(), where + " after compute")
        if k:
            assert cpu.make_landmarks(2) == gpu.make_landmarks(2)
    cpu.close()
    gpu.close()
