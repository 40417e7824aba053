"""Oracle aligners: C restatement vs an independent numpy restatement of the reference formulas
(stereouv_aligner.cpp:72-187, uvd_aligner.cpp:77-171), analytic Jacobian vs central differences under
the v2t update, 6x6 solve, and convergence to the known synthetic motion."""
import numpy as np
import pytest

from oracle import tier_a
from vslam_b200 import synth

T0 = np.hstack([np.eye(3), np.zeros((3, 1))])


def _problem(kind, n, kernel=16.0, seed=424242):
    cam = synth.camera("kitti")
    c = synth.correspondences(n, kind, cam, seed=seed)
    omega = c["omega"] if kind == "stereouv" else np.stack([c["omega_uv"], c["omega_d"]], 1)
    al = tier_a.Aligner(kind, c["moving"], c["fixed"], omega, c["wt"], cam.K, cam.baseline, cam.rows, cam.cols,
                        0.1, kernel)
    return cam, c, omega, al


def _skew(p):
    return np.array([[0, -p[2], p[1]], [p[2], 0, -p[0]], [-p[1], p[0], 0]])


def _numpy_linearize(kind, cam, c, omega, T, kernel, ignore_outliers, min_depth=0.1):
    """Matrix-form restatement with numpy (independent of the C code's scalar expansion)."""
    n = len(c["moving"])
    H, b, total, inl = np.zeros((6, 6)), np.zeros(6), 0.0, 0
    errors, inliers = np.full(n, -1.0), np.zeros(n, bool)
    K, base = cam.K, cam.baseline
    for u in range(n):
        p = T[:, :3] @ c["moving"][u] + T[:, 3]
        if kind == "stereouv":
            if p[2] < min_depth:
                continue
            abc_l = K @ p
            abc_r = abc_l + base
            il, ir = abc_l / abc_l[2], abc_r / abc_r[2]
            if il[0] < 0 or il[0] > cam.cols or il[1] < 0 or il[1] > cam.rows:
                continue
            if ir[0] < 0 or ir[0] > cam.cols or ir[1] < 0 or ir[1] > cam.rows:
                continue
            e = np.array([il[0], il[1], ir[0], ir[1]]) - c["fixed"][u]
            Om = np.eye(4) * omega[u]
        else:
            if p[2] <= min_depth:
                continue
            uvd = K @ p
            pi = uvd / uvd[2]
            if pi[0] < 0 or pi[0] > cam.cols or pi[1] < 0 or pi[1] > cam.rows:
                continue
            e = np.array([pi[0], pi[1], p[2]]) - c["fixed"][u]
            Om = np.diag([omega[u, 0], omega[u, 0], omega[u, 1]])
        chi = e @ Om @ e
        errors[u] = chi
        if chi > kernel:
            if ignore_outliers:
                continue
            Om = Om * (kernel / chi)
        else:
            inliers[u] = True
            inl += 1
        total += chi
        Jt = np.hstack([c["wt"][u] * np.eye(3), -2 * _skew(p)])
        if kind == "stereouv":
            KJ = K @ Jt
            Jl = np.array([[1 / abc_l[2], 0, -abc_l[0] / abc_l[2] ** 2], [0, 1 / abc_l[2], -abc_l[1] / abc_l[2] ** 2]])
            Jr = np.array([[1 / abc_r[2], 0, -abc_r[0] / abc_r[2] ** 2], [0, 1 / abc_r[2], -abc_r[1] / abc_r[2] ** 2]])
            J = np.vstack([Jl @ KJ, Jr @ KJ])
        else:
            Jp = np.array([[1 / p[2], 0, -uvd[0] / p[2] ** 2], [0, 1 / p[2], -uvd[1] / p[2] ** 2], [0, 0, 1]])
            J = Jp @ K @ Jt
        H += J.T @ Om @ J
        b += J.T @ Om @ e
    return H, b, total, inl, errors, inliers


@pytest.mark.parametrize("kind", ["stereouv", "uvd"])
@pytest.mark.parametrize("ignore", [False, True])
def test_linearize_matches_numpy_restatement(kind, ignore):
    cam, c, omega, al = _problem(kind, 600, kernel=16.0)
    T = synth.true_motion() * 1.0
    T[:, 3] += [0.02, -0.01, 0.05]                      # near the optimum: mix of inliers and outliers
    s = al.linearize(T, ignore)
    H, b, total, inl, errors, inliers = _numpy_linearize(kind, cam, c, omega, T, 16.0, ignore)
    assert s["inliers"] == inl and s["outliers"] == 600 - inl and 50 < inl < 600
    np.testing.assert_allclose(s["H"], H, rtol=1e-10, atol=1e-6)
    np.testing.assert_allclose(s["b"], b, rtol=1e-10, atol=1e-6)
    np.testing.assert_allclose(s["total_error"], total, rtol=1e-12)
    np.testing.assert_allclose(al.errors, errors, rtol=1e-12)
    assert np.array_equal(al.inliers.astype(bool), inliers)


def test_skip_rules_leave_minus_one():
    cam = synth.camera("kitti")
    moving = np.array([[0.0, 0.0, 0.05], [0.0, 0.0, 10.0], [100.0, 0.0, 10.0], [0.0, 0.0, 0.1]])
    fixed = np.tile([607.0, 185.0, 568.0, 185.0], (4, 1))
    al = tier_a.Aligner("stereouv", moving, fixed, np.ones(4), np.ones(4), cam.K, cam.baseline, cam.rows, cam.cols,
                        0.1, 1e9)
    s = al.linearize(T0)
    assert list(al.errors < 0) == [True, False, True, True]   # depth<min, ok, outside FOV, z == min_depth ok for
    # StereoUV (strict <) but its right projection leaves the image
    al2 = tier_a.Aligner("uvd", moving, fixed[:, :3], np.ones((4, 2)), np.ones(4), cam.K, cam.baseline, cam.rows,
                         cam.cols, 0.1, 1e9)
    al2.linearize(T0)
    assert list(al2.errors < 0) == [True, False, True, True]  # UVD: depth <= min_depth skipped (uvd_aligner.cpp:95)
    assert s["inliers"] == 1 and s["outliers"] == 3


@pytest.mark.parametrize("kind", ["stereouv", "uvd"])
def test_analytic_jacobian_matches_central_differences_under_v2t(kind):
    """b = sum J^T Omega e must be the gradient of 0.5*chi2 w.r.t. the v2t perturbation (wt = 1)."""
    cam = synth.camera("kitti")
    c = synth.correspondences(50, kind, cam, seed=7, outlier_fraction=0.0)
    omega = np.ones(50) if kind == "stereouv" else np.ones((50, 2))
    al = tier_a.Aligner(kind, c["moving"], c["fixed"], omega, np.ones(50), cam.K, cam.baseline, cam.rows, cam.cols,
                        0.1, 1e12)
    T = synth.true_motion()
    s = al.linearize(T)

    def chi2(dx):
        D = tier_a.v2t(dx)
        Tn = np.hstack([D[:, :3] @ T[:, :3], (D[:, :3] @ T[:, 3] + D[:, 3])[:, None]])
        return al.linearize(Tn)["total_error"]

    g = np.zeros(6)
    for i in range(6):
        h = 1e-6
        d = np.zeros(6)
        d[i] = h
        g[i] = (chi2(d) - chi2(-d)) / (2 * h)
    np.testing.assert_allclose(0.5 * g, s["b"], rtol=2e-5, atol=1e-4)


def test_solve6_and_v2t():
    rng = np.random.default_rng(3)
    A = rng.normal(size=(6, 6))
    A = A @ A.T + np.eye(6)
    b = rng.normal(size=6)
    np.testing.assert_allclose(tier_a.solve6(A, b), np.linalg.solve(A, b), rtol=1e-11)
    T = tier_a.v2t([1, 2, 3, 0.01, -0.02, 0.03])
    np.testing.assert_allclose(T[:, :3] @ T[:, :3].T, np.eye(3), atol=1e-15)
    np.testing.assert_allclose(T[:, 3], [1, 2, 3])
    # small-angle: rotation vector ~ 2*q
    np.testing.assert_allclose(T[:, :3], np.eye(3) + 2 * _skew([0.01, -0.02, 0.03]), atol=3e-3)
    T = tier_a.v2t([0, 0, 0, 2.0, 0, 0])               # |q|^2 >= 1 branch: w = 0, q normalised -> 180 deg about x
    np.testing.assert_allclose(T[:, :3], np.diag([1.0, -1.0, -1.0]), atol=1e-15)


@pytest.mark.parametrize("kind,damping,kernel", [("stereouv", 0.0, 16.0), ("stereouv", 5.0, 4.0), ("uvd", 0.0, 16.0)])
def test_converge_recovers_true_motion(kind, damping, kernel):
    cam, c, omega, al = _problem(kind, 20000, kernel=kernel)
    r = al.converge(T0, damping, 1e-3, 1000, 0 if damping == 0 else 100)
    assert r["converged"]
    Tt = c["T_true"]
    dR = r["T"][:, :3] @ Tt[:, :3].T
    ang = np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1))
    assert ang < 2e-4 and np.linalg.norm(r["T"][:, 3] - Tt[:, 3]) < 5e-3
    assert r["inliers"] > 0.6 * 20000
    np.testing.assert_allclose(r["T"][:, :3] @ r["T"][:, :3].T, np.eye(3), atol=1e-9)
