"""Host-side wire formats of SURVEY 8f row 4 through the C ABI (no GPU needed): the KITTI / TUM trajectory lines of
WorldMap::writeTrajectoryKITTI / writeTrajectoryTUM (reference src/types/world_map.cpp:183-252) and the 3x3
complete-pivoting solve, against the oracle."""
import numpy as np

from oracle import tier_a
from vslam_b200 import api, synth


def _poses(n, seed=3):
    h = synth.landmark_histories(2, n_frames=n, seed=seed)
    return h["camera_to_world"]


def test_trajectory_lines_equal_the_oracle_and_the_stream_format():
    T = _poses(25)
    ts = 1403636579.763555527 + 0.05 * np.arange(len(T))
    for i in range(len(T)):
        assert api.format_trajectory(T[i]) == tier_a.format_trajectory(T[i])
        assert api.format_trajectory(T[i], ts[i]) == tier_a.format_trajectory(T[i], ts[i])
    line = api.format_trajectory(T[3])
    assert line.endswith(" \n") and len(line.split()) == 12 and all(len(v.split(".")[1]) == 9 for v in line.split())
    assert len(api.format_trajectory(T[3], ts[3]).split()) == 8


def test_trajectory_files(tmp_path):
    T = _poses(40, seed=9)
    ts = 100.0 + 0.1 * np.arange(len(T))
    api.write_trajectory(tmp_path / "kitti.txt", T)
    api.write_trajectory(tmp_path / "tum.txt", T, ts)
    kitti = (tmp_path / "kitti.txt").read_text()
    assert kitti == "".join(tier_a.format_trajectory(t) for t in T)
    back = np.loadtxt(tmp_path / "kitti.txt")                      # what KITTI's evaluation tools read
    np.testing.assert_allclose(back, T, atol=5.1e-10)
    tum = np.loadtxt(tmp_path / "tum.txt")
    assert tum.shape == (len(T), 8)
    np.testing.assert_allclose(tum[:, 0], ts, atol=1e-9)
    np.testing.assert_allclose(tum[:, 1:4], T.reshape(-1, 3, 4)[:, :, 3], atol=5.1e-10)
    from scipy.spatial.transform import Rotation
    want = Rotation.from_matrix(T.reshape(-1, 3, 4)[:, :, :3]).as_quat()
    sign = np.sign(np.sum(want * tum[:, 4:], axis=1))[:, None]
    np.testing.assert_allclose(tum[:, 4:] * sign, want, atol=1e-8)
    api.write_trajectory(tmp_path / "kitti.txt", T[:2])            # overwriting, as the reference
    assert len((tmp_path / "kitti.txt").read_text().splitlines()) == 2


def test_solve3_equals_the_oracle_bit_for_bit():
    rng = np.random.default_rng(1)
    for _ in range(100):
        A, b = rng.normal(size=(3, 3)), rng.normal(size=3)
        assert np.array_equal(api.solve3(A, b), tier_a.solve3(A, b))
    assert np.array_equal(api.solve3(np.zeros((3, 3)), np.ones(3)), np.zeros(3))
