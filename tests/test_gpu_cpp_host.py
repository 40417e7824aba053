"""The C++14 host layer (include/vslam_b200.hpp) used from a compiled C++ program shaped like the reference's
executables/test_stereo_frontend.cpp, checked against the CPU oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import pipeline, tier_a
from vslam_b200 import api, configs, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_api_check.cpp")
LIBDIR = os.path.dirname(api.LIB_PATH)


def _build(out):
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), SRC, "-o", out,
                           "-L", LIBDIR, "-lvslam_b200", "-Wl,-rpath," + LIBDIR])


def test_cpp_host_layer_compiles_as_cxx14_and_header_as_c99(tmp_path):
    _build(str(tmp_path / "host_api_check"))
    c = tmp_path / "abi.c"
    c.write_text('#include "vslam_b200.h"\nint main(void) { return vslam_device_count() < 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(c),
                           "-o", str(tmp_path / "abi"), "-L", LIBDIR, "-lvslam_b200", "-Wl,-rpath," + LIBDIR])
    subprocess.check_call([str(tmp_path / "abi")])


@pytest.mark.gpu
def test_cpp_frontend_program_matches_oracle(tmp_path):
    exe = str(tmp_path / "host_api_check")
    _build(exe)
    cfg, cam = configs.KITTI, synth.camera("kitti")
    world = synth.BandWorld(cam.cols, cam.rows, 12, max_frames=4)
    left, right = world.pair(0)
    left1, right1 = world.pair(1)
    left.tofile(tmp_path / "left.u8")
    right.tofile(tmp_path / "right.u8")
    left1.tofile(tmp_path / "left1.u8")
    right1.tofile(tmp_path / "right1.u8")
    n = 3000
    c = synth.correspondences(n, "stereouv", cam, seed=99)
    np.concatenate([[float(n)], c["moving"].ravel(), c["fixed"].ravel(), c["omega"], c["wt"]]).tofile(
        tmp_path / "correspondences.f64")
    hist = synth.landmark_histories(200, n_frames=30, seed=17, outlier_fraction=0.05)
    with open(tmp_path / "landmarks.bin", "wb") as f:
        nl, nm, nf = 200, int(hist["offsets"][-1]), 30
        f.write(np.array([nl, nm, nf, 0], np.int32).tobytes())
        f.write(hist["offsets"].astype(np.int32).tobytes())
        if (nl + 1) % 2:
            f.write(b"\0" * 4)                      # keep the 8-byte fields aligned
        f.write(hist["measurements"].tobytes())
        f.write(hist["world_to_camera"].tobytes())
        f.write(hist["camera_to_world"].tobytes())
        f.write(hist["world"].tobytes())
        f.write(hist["number_of_updates"].astype(np.uint32).tobytes())
    out = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    lines = dict(l.split(" ", 1) for l in out.stdout.strip().splitlines())

    o = pipeline.StereoFramePointGeneratorOracle(cfg, cam, "a").initialize(left, right, True)
    o.compute()
    fp = o.framepoints()
    assert lines["features"] == "%d %d" % (len(o.kps_left), len(o.kps_right))
    assert lines["points"] == "%d new %d threshold %.1f" % (len(fp), len(o.matches), o.thresholds.mean())
    h = 1469598103934665603
    for b in o.desc_left.ravel().tolist():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    for p in fp:
        for x in (p["index_left"], p["index_right"], p["distance"], p["epipolar_offset"]):
            h = ((h ^ (int(x) & 0xFFFFFFFF)) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    assert lines["hash"] == str(h)
    assert [float(v) for v in lines["first"].split()] == fp[0]["cam"].tolist()

    # second frame: track -> compute -> recoverPoints
    from test_oracle_track import previous_points
    prev = previous_points(o, fp)
    o.initialize(left1, right1, False)
    T = np.hstack([np.eye(3), np.zeros((3, 1))])
    T[0, 3] = -(386.1448 / 718.856) / 4
    r = o.track(prev, T, False, 15, 25.6)
    o.compute(o.tracked_points(r["tracks"]))
    new = o.framepoints()
    h = 1469598103934665603
    for t in r["tracks"]:
        for x in (t["index_previous"], t["index_left"], t["index_right"], t["distance"], t["epipolar_offset"]):
            h = ((h ^ (int(x) & 0xFFFFFFFF)) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    for x in r["lost"].tolist() + new["index_left"].tolist():
        h = ((h ^ (int(x) & 0xFFFFFFFF)) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    f = lines["tracking"].split()
    assert f[:7] == [str(len(r["tracks"])), "lost", str(len(r["lost"])), "landmarks", str(r["tracked_landmarks"]), "new",
                     str(len(new))]
    assert float(f[8]) == r["accumulated_distance"] / len(r["tracks"]) and f[10] == str(h)
    assert len(r["tracks"]) > 300
    rec = o.recover_points(prev[r["lost"]], T, 64.0)
    h = 1469598103934665603
    for q in rec:
        h = ((h ^ (int(q["index_lost"]) & 0xFFFFFFFF)) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        for b in q["desc_left"].tolist():
            h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    assert lines["recovered"] == "%d hash %d" % (len(rec), h)

    al = tier_a.Aligner("stereouv", c["moving"], c["fixed"], c["omega"], c["wt"], cam.K, cam.baseline, cam.rows, cam.cols,
                        0.1, 16.0)
    r = al.converge(np.hstack([np.eye(3), np.zeros((3, 1))]), 0.0, 1e-3, 1000, 0)
    assert lines["aligner"] == "converged %d rounds %d inliers %d outliers %d" % (r["converged"], r["rounds"], r["inliers"],
                                                                                   r["outliers"])
    T = np.array([float(v) for v in lines["pose"].split()]).reshape(3, 4)
    dR = T[:, :3] @ r["T"][:, :3].T
    assert np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1)) <= 1e-6 and np.linalg.norm(T[:, 3] - r["T"][:, 3]) <= 1e-5
    assert "called with empty frame" in lines["exception"]

    # landmark refinement and trajectory files through the C++ classes
    h = 1469598103934665603
    world, updates, outcomes = hist["world"].copy(), hist["number_of_updates"].copy(), []
    for i in range(200):
        ms = hist["measurements"][hist["offsets"][i]:hist["offsets"][i + 1]]
        world[i], updates[i], oc, _ = tier_a.landmark_update(ms, hist["world_to_camera"], hist["camera_to_world"],
                                                             hist["world"][i], hist["number_of_updates"][i])
        outcomes.append(oc)
    for bits in world.ravel().view(np.uint64).tolist():
        h = ((h ^ bits) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    for x in updates.tolist() + outcomes:
        h = ((h ^ int(x)) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    assert lines["landmarks"] == "200 hash %d" % h
    assert lines["trajectories"] == "30"
    ts = 1403636579.763555527 + 0.05 * np.arange(30)
    assert (tmp_path / "trajectory_kitti.txt").read_text() == "".join(tier_a.format_trajectory(t) for t in hist["camera_to_world"])
    assert (tmp_path / "trajectory_tum.txt").read_text() == "".join(
        tier_a.format_trajectory(t, s) for t, s in zip(hist["camera_to_world"], ts))
