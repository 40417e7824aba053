"""Oracle of SURVEY 8f row 4 -- Landmark::update (src/types/landmark.cpp:66-152) and the trajectory wire formats
(src/types/world_map.cpp:183-252) -- pinned against independent restatements: a numpy matrix-form Gauss-Newton with
numpy.linalg.solve, scipy's rotation -> quaternion, and Python's own "%.9f" formatting of the reference's stream
manipulators (std::fixed, setprecision(9)).  The reference holds no fixture for either (parity unpinned beyond this)."""
import numpy as np
import pytest

from oracle import tier_a
from vslam_b200 import synth


def _numpy_landmark_update(ms, w2c, c2w, world, n_updates, max_iterations=100, max_err2=25.0):
    """literal matrix-form restatement of landmark.cpp:82-167"""
    x = np.array(world, np.float64)
    prev = 0.0
    for it in range(max_iterations):
        H, b, total, outliers = np.zeros((3, 3)), np.zeros(3), 0.0, 0
        for m in ms:
            W = w2c[m["frame"]].reshape(3, 4)
            p = W[:, :3] @ x + W[:, 3]
            if p[2] <= 0:
                outliers += 1
                continue
            e = p - m["camera_coordinates"]
            omega = np.eye(3) * m["inverse_depth_meters"]
            err2 = float(e @ omega @ e)
            total += err2
            if err2 > max_err2:
                omega = omega * (max_err2 / err2)
                outliers += 1
            J = W[:, :3]
            H += J.T @ omega @ J
            b += J.T @ omega @ e
        # (every measurement skipped: H = 0; Eigen's full-pivot LU has rank 0 and solve() returns 0)
        x = x + (np.linalg.solve(H, -b) if np.any(H) else np.zeros(3))
        if abs(total - prev) < 1e-5 or it == 999:
            inliers = len(ms) - outliers
            if inliers > n_updates:
                return x, inliers, 1, it + 1
            if inliers < outliers:
                acc = np.zeros(3)
                for m in ms:
                    C = c2w[m["frame"]].reshape(3, 4)
                    acc += C[:, :3] @ m["camera_coordinates"] + C[:, 3]
                return acc / len(ms), n_updates, 2, it + 1
            return np.array(world, np.float64), n_updates, 3, it + 1
        prev = total
    return np.array(world, np.float64), n_updates, 0, max_iterations


@pytest.mark.parametrize("seed,outliers", [(1, 0.0), (2, 0.05), (3, 0.3)])
def test_landmark_update_matches_the_matrix_form(seed, outliers):
    h = synth.landmark_histories(60, n_frames=30, seed=seed, outlier_fraction=outliers)
    outcomes = set()
    for i in range(60):
        ms = h["measurements"][h["offsets"][i]:h["offsets"][i + 1]]
        got = tier_a.landmark_update(ms, h["world_to_camera"], h["camera_to_world"], h["world"][i], h["number_of_updates"][i])
        want = _numpy_landmark_update(ms, h["world_to_camera"], h["camera_to_world"], h["world"][i],
                                      int(h["number_of_updates"][i]))
        assert got[1:] == want[1:], (i, got, want)
        np.testing.assert_allclose(got[0], want[0], rtol=1e-10, atol=1e-9)
        outcomes.add(got[2])
    assert 1 in outcomes


def test_landmark_update_converges_to_the_observed_point():
    h = synth.landmark_histories(40, n_frames=40, seed=11, noise=0.002, outlier_fraction=0.0)
    for i in range(40):
        ms = h["measurements"][h["offsets"][i]:h["offsets"][i + 1]]
        x, n_up, outcome, _ = tier_a.landmark_update(ms, h["world_to_camera"], h["camera_to_world"], h["world"][i], 0)
        assert outcome == 1 and n_up == len(ms)
        assert np.linalg.norm(x - h["truth"][i]) < 0.05 * np.linalg.norm(ms["camera_coordinates"][-1])


def test_landmark_update_branches():
    h = synth.landmark_histories(8, n_frames=12, seed=5, outlier_fraction=0.0)
    ms = h["measurements"][h["offsets"][0]:h["offsets"][1]]
    w2c, c2w = h["world_to_camera"], h["camera_to_world"]
    # more updates recorded than inliers now: converged, state kept (landmark.cpp:143 false, :150 false)
    x, n_up, outcome, _ = tier_a.landmark_update(ms, w2c, c2w, h["world"][0], 1000)
    assert outcome == 3 and n_up == 1000 and np.array_equal(x, h["world"][0])
    # iteration cap reached before the error settles: nothing is written
    x, n_up, outcome, it = tier_a.landmark_update(ms, w2c, c2w, h["world"][0] + 3.0, 0, max_iterations=1)
    assert outcome == 0 and it == 1 and n_up == 0 and np.array_equal(x, h["world"][0] + 3.0)
    # an estimate behind every camera: all measurements are skipped as outliers (:103-106); H = 0, full-pivot LU of a
    # zero matrix has rank 0 and solve() returns 0 -> error 0 twice -> converged with inliers < outliers -> average reset
    far = h["world"][0] - np.array([0.0, 0.0, 500.0])
    x, n_up, outcome, it = tier_a.landmark_update(ms, w2c, c2w, far, 3)
    want = np.mean([c2w[m["frame"]].reshape(3, 4)[:, :3] @ m["camera_coordinates"] + c2w[m["frame"]].reshape(3, 4)[:, 3]
                    for m in ms], axis=0)
    assert outcome == 2 and n_up == 3 and it == 1
    np.testing.assert_allclose(x, want, rtol=1e-12)


def test_solve3_and_quaternion():
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(0)
    for _ in range(50):
        A, b = rng.normal(size=(3, 3)), rng.normal(size=3)
        np.testing.assert_allclose(tier_a.solve3(A, b), np.linalg.solve(A, b), rtol=1e-9, atol=1e-12)
    assert np.array_equal(tier_a.solve3(np.zeros((3, 3)), np.ones(3)), np.zeros(3))
    for rv in list(rng.normal(size=(50, 3))) + [np.array([np.pi, 0, 0]), np.array([0, np.pi - 1e-9, 0]), np.zeros(3),
                                                 np.array([0, 0, 3.0])]:
        R = Rotation.from_rotvec(rv).as_matrix()
        q = tier_a.rotation_to_quaternion(R)
        want = Rotation.from_matrix(R).as_quat()          # (x, y, z, w), sign-ambiguous
        assert min(np.abs(q - want).max(), np.abs(q + want).max()) < 1e-9
        assert abs(np.linalg.norm(q) - 1) < 1e-12


def test_trajectory_lines_follow_the_reference_stream_format():
    T = np.array([[0.9999, -0.01, 0.002, 12.3456789012], [0.01, 0.9999, 0.0, -0.5], [-0.002, 0.0, 1.0, 1e-10]])
    kitti = tier_a.format_trajectory(T)
    assert kitti == "".join("%.9f " % v for v in T.reshape(12)) + "\n"           # world_map.cpp:196-214
    assert kitti.startswith("0.999900000 -0.010000000 0.002000000 12.345678901 ")
    tum = tier_a.format_trajectory(T, timestamp=1403636579.763555527)
    q = tier_a.rotation_to_quaternion(T[:, :3])
    want = "%.9f " % 1403636579.763555527 + "".join("%.9f " % v for v in (T[0, 3], T[1, 3], T[2, 3], *q)) + "\n"
    assert tum == want                                                          # world_map.cpp:230-248
