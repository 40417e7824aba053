"""Parity of the CUDA aligner linearisation (through the C ABI) with the CPU oracle.
Tolerances (north_star / SURVEY 8c): errors[] and inlier sets bit-exact (same expression order, no contraction);
H, b, total error 1e-10 relative (summation order differs); converged pose <= 1e-6 rad, <= 1e-5 m."""
import numpy as np
import pytest

from oracle import tier_a
from vslam_b200 import api, configs, synth

pytestmark = pytest.mark.gpu

T0 = np.hstack([np.eye(3), np.zeros((3, 1))])


def _pair(kind, n, acfg, seed=424242, cam=None):
    cam = cam or synth.camera("kitti")
    c = synth.correspondences(n, kind, cam, seed=seed)
    omega = c["omega"] if kind == "stereouv" else np.stack([c["omega_uv"], c["omega_d"]], 1)
    cls = api.StereoUVAligner if kind == "stereouv" else api.UVDAligner
    gpu = cls(acfg, max_points=max(n, 1))
    gpu.initialize(c["moving"], c["fixed"], omega, c["wt"], cam.K, cam.baseline, cam.rows, cam.cols, T0)
    cpu = tier_a.Aligner(kind, c["moving"], c["fixed"], omega, c["wt"], cam.K, cam.baseline, cam.rows, cam.cols,
                         acfg.minimum_reliable_depth_meters, acfg.maximum_error_kernel)
    return gpu, cpu, c


def _close(got, want):
    scale = np.abs(want["H"]).max()
    np.testing.assert_allclose(got["H"], want["H"], rtol=1e-10, atol=1e-12 * scale)
    # b is a cancelling sum near the optimum: bound the absolute error by the Cauchy-Schwarz norm of its terms,
    # sum |J_i w e| <= sqrt(H_ii * chi2), at 1e-13 of that norm
    b_atol = 1e-13 * np.sqrt(np.abs(np.diag(want["H"])) * max(want["total_error"], 0.0))
    assert np.all(np.abs(got["b"] - want["b"]) <= 1e-10 * np.abs(want["b"]) + b_atol), (got["b"], want["b"])
    np.testing.assert_allclose(got["total_error"], want["total_error"], rtol=1e-12)
    assert got["inliers"] == want["inliers"] and got["outliers"] == want["outliers"]


def _pose_delta(A, B):
    dR = A[:, :3] @ B[:, :3].T
    return np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1)), np.linalg.norm(A[:, 3] - B[:, 3])


@pytest.mark.parametrize("kind", ["stereouv", "uvd"])
@pytest.mark.parametrize("ignore", [False, True])
@pytest.mark.parametrize("n", [1, 37, 5000, 100000])
def test_linearize_matches_oracle(kind, ignore, n):
    gpu, cpu, c = _pair(kind, n, configs.KITTI_FAST_ALIGNER)
    T = synth.true_motion().copy()
    T[:, 3] += [0.02, -0.01, 0.05]
    gpu.setPreviousToCurrent(T)
    got, want = gpu.linearize(ignore), cpu.linearize(T, ignore)
    _close(got, want)
    assert np.array_equal(gpu.errors(), cpu.errors)                  # bit-exact chi per correspondence
    assert np.array_equal(gpu.inliers(), cpu.inliers.astype(bool))
    assert np.allclose(got["H"], got["H"].T)
    gpu.close()


def test_empty_problem():
    gpu = api.StereoUVAligner(configs.KITTI_ALIGNER, max_points=16)
    cam = synth.camera("kitti")
    gpu.initialize(np.zeros((0, 3)), np.zeros((0, 4)), np.zeros(0), np.zeros(0), cam.K, cam.baseline, cam.rows, cam.cols)
    s = gpu.linearize()
    assert s["inliers"] == 0 and s["outliers"] == 0 and s["total_error"] == 0 and not s["H"].any()
    gpu.close()


def test_linearize_is_run_to_run_deterministic():
    gpu, _, _ = _pair("stereouv", 100000, configs.KITTI_ALIGNER)
    a = gpu.linearize()
    for _ in range(5):
        b = gpu.linearize()
        assert np.array_equal(a["H"], b["H"]) and np.array_equal(a["b"], b["b"]) and a["total_error"] == b["total_error"]
    gpu.close()


@pytest.mark.parametrize("kind,acfg", [("stereouv", configs.KITTI_FAST_ALIGNER), ("stereouv", configs.KITTI_ALIGNER),
                                       ("uvd", configs.KITTI_FAST_ALIGNER)])
def test_config4_stress_ten_rounds(kind, acfg):
    """BASELINE.json config 4: 100k correspondences, exactly 10 oneRound(false)."""
    gpu, cpu, c = _pair(kind, 100000, acfg)
    T = T0.copy()
    for _ in range(10):
        got = gpu.oneRound(False)
        T, want = cpu.one_round(T, acfg.damping, False)
        _close(got, want)
    ang, dist = _pose_delta(gpu.previousToCurrent(), T)
    assert ang <= 1e-6 and dist <= 1e-5
    gpu.close()


@pytest.mark.parametrize("kind,acfg", [("stereouv", configs.KITTI_FAST_ALIGNER), ("stereouv", configs.EUROC_ALIGNER),
                                       ("uvd", configs.KITTI_FAST_ALIGNER)])
def test_converge_matches_oracle(kind, acfg):
    gpu, cpu, c = _pair(kind, 20000, acfg)
    got = gpu.converge()
    want = cpu.converge(T0, acfg.damping, acfg.error_delta_for_convergence, acfg.maximum_number_of_iterations,
                        acfg.minimum_number_of_inliers)
    assert gpu.has_system_converged == want["converged"] and gpu.number_of_rounds == want["rounds"]
    ang, dist = _pose_delta(gpu.previousToCurrent(), want["T"])
    assert ang <= 1e-6 and dist <= 1e-5
    assert got["inliers"] == want["inliers"]
    np.testing.assert_allclose(gpu.information_matrix, want["information"], rtol=1e-8,
                               atol=1e-12 * np.abs(want["information"]).max())
    assert np.array_equal(gpu.inliers(), cpu.inliers.astype(bool))
    # and it is the right answer
    ang, dist = _pose_delta(gpu.previousToCurrent(), c["T_true"])
    assert ang < 5e-4 and dist < 2e-2
    gpu.close()


def test_linearity_over_disjoint_halves_at_full_size():
    """size-independent property at 2M correspondences: H, b, chi2, inliers of a set = sum over a partition."""
    n = 2_000_000
    cam = synth.camera("kitti")
    c = synth.correspondences(n, "stereouv", cam, seed=5)
    acfg = configs.KITTI_ALIGNER
    T = synth.true_motion()

    def run(sl):
        g = api.StereoUVAligner(acfg, max_points=n)
        g.initialize(c["moving"][sl], c["fixed"][sl], c["omega"][sl], c["wt"][sl], cam.K, cam.baseline, cam.rows,
                     cam.cols, T)
        s = g.linearize()
        g.close()
        return s

    whole, a, b = run(slice(None)), run(slice(0, 700_001)), run(slice(700_001, None))
    np.testing.assert_allclose(whole["H"], a["H"] + b["H"], rtol=1e-11)
    np.testing.assert_allclose(whole["b"], a["b"] + b["b"], rtol=1e-9, atol=1e-6)
    np.testing.assert_allclose(whole["total_error"], a["total_error"] + b["total_error"], rtol=1e-12)
    assert whole["inliers"] == a["inliers"] + b["inliers"]


def test_capacity_error():
    gpu = api.UVDAligner(configs.KITTI_ALIGNER, max_points=64)
    cam = synth.camera("kitti")
    with pytest.raises(api.VslamError) as e:
        gpu.initialize(np.zeros((200, 3)), np.zeros((200, 3)), np.zeros((200, 2)), np.zeros(200), cam.K, cam.baseline,
                       cam.rows, cam.cols)
    assert e.value.code == -3
    gpu.close()


@pytest.mark.parametrize("kind,acfg,n", [("stereouv", configs.KITTI_FAST_ALIGNER, 20000), ("stereouv", configs.KITTI_ALIGNER, 700),
                                         ("stereouv", configs.EUROC_ALIGNER, 100000), ("uvd", configs.KITTI_FAST_ALIGNER, 20000),
                                         ("uvd", configs.KITTI_ALIGNER, 3)])
def test_fused_gauss_newton_is_bit_identical_to_the_stepwise_driver(kind, acfg, n):
    """vslam_aligner_converge_fused (one persistent cooperative kernel) vs vslam_aligner_converge (host loop)."""
    gpu, cpu, c = _pair(kind, n, acfg)
    a = gpu.converge(fused=False)
    Ta, ra, ca, ia = gpu.previousToCurrent(), gpu.number_of_rounds, gpu.has_system_converged, gpu.information_matrix.copy()
    ea, ina = gpu.errors(), gpu.inliers()
    gpu.setPreviousToCurrent(T0)
    b = gpu.converge(fused=True)
    assert gpu.number_of_rounds == ra and gpu.has_system_converged == ca and ra >= 1
    assert np.array_equal(gpu.previousToCurrent(), Ta)
    assert np.array_equal(a["H"], b["H"]) and np.array_equal(a["b"], b["b"]) and a["total_error"] == b["total_error"]
    assert a["inliers"] == b["inliers"] and np.array_equal(gpu.information_matrix, ia)
    assert np.array_equal(gpu.errors(), ea) and np.array_equal(gpu.inliers(), ina)
    gpu.close()


def test_fused_gauss_newton_iteration_cap():
    import dataclasses
    acfg = dataclasses.replace(configs.KITTI_FAST_ALIGNER, maximum_number_of_iterations=3, error_delta_for_convergence=1e-12)
    gpu, cpu, c = _pair("stereouv", 5000, acfg)
    gpu.converge(fused=True)
    want = cpu.converge(T0, acfg.damping, acfg.error_delta_for_convergence, 3, acfg.minimum_number_of_inliers)
    assert not gpu.has_system_converged and not want["converged"]
    assert gpu.number_of_rounds == want["rounds"] == 3
    gpu.close()
