"""Seeded synthetic inputs for the hot path.

Nothing here is reference code: the reference ships no datasets; these
generators define the workloads of BASELINE.json's configs so that the CPU oracle and the CUDA path
consume byte-identical inputs.

* band_world(): rectified stereo pairs / sequences of KITTI (1241x376), EuRoC (752x480) or HD
  (1920x1080) shape.  Horizontal bands at integer disparities, so every true match lies on the
  same image row (epipolar offset 0) and the ground-truth depth of band b is -b_x / d_b.
* correspondences(): the aligner stress set (config 4): 3D-2D correspondences with noise and gross
  outliers for StereoUVAligner / UVDAligner.
"""
from __future__ import annotations

import dataclasses

import numpy as np

# name -> (cols, rows, fx, fy, cx, cy, b_x)   b_x = baselineHomogeneous()(0) < 0
# (/root/reference/src/framepoint_generation/stereo_framepoint_generator.cpp:26-28)
CAMERAS = {
    "kitti": (1241, 376, 718.856, 718.856, 607.1928, 185.2157, -386.1448),
    "euroc": (752, 480, 435.2, 435.2, 367.4, 252.2, -47.9),
    "hd": (1920, 1080, 1400.0, 1400.0, 960.0, 540.0, -168.0),
}


@dataclasses.dataclass
class Camera:
    cols: int
    rows: int
    fx: float
    fy: float
    cx: float
    cy: float
    bx: float

    @property
    def K(self) -> np.ndarray:
        return np.array([[self.fx, 0.0, self.cx], [0.0, self.fy, self.cy], [0.0, 0.0, 1.0]])

    @property
    def baseline(self) -> np.ndarray:
        return np.array([self.bx, 0.0, 0.0])


def camera(name: str) -> Camera:
    return Camera(*CAMERAS[name])


BAND_ROWS = 47
_RHO = 900.0 / (376.0 * 1241.0)


class BandWorld:
    """One static 3-D consistent world; frame k is the view after a camera x-translation of k*B/4."""

    def __init__(self, cols: int, rows: int, seed: int, max_frames: int = 1):
        import cv2  # data generation only
        self.cols, self.rows, self.seed = cols, rows, seed
        margin = max(1024, 12 * max_frames + 64)
        rng = np.random.default_rng(seed)
        wc = cols + margin
        canvas = np.full((rows, wc), 96.0, dtype=np.float32)
        self.band_disparity = []
        self.band_start = list(range(0, rows, BAND_ROWS))
        for y0 in self.band_start:
            h = min(BAND_ROWS, rows - y0)
            self.band_disparity.append(4 * int(rng.integers(1, 13)))
            if h < 7:
                continue
            n = int(1.6 * _RHO * h * wc)
            ws = rng.integers(6, 40, n)
            hs = rng.integers(6, min(40, h), n) if min(40, h) > 6 else np.full(n, 6)
            xs = (rng.random(n) * (wc - ws)).astype(np.int64)
            ys = y0 + (rng.random(n) * (h - hs + 1)).astype(np.int64)
            vs = rng.integers(20, 236, n)
            for i in range(n):
                canvas[ys[i]:ys[i] + hs[i], xs[i]:xs[i] + ws[i]] = vs[i]
        self.canvas = cv2.GaussianBlur(canvas, (0, 0), 0.8, borderType=cv2.BORDER_REPLICATE)

    def pair(self, k: int = 0, noise_seed: int | None = None, sigma: float = 1.5):
        """-> (left, right) uint8 [rows, cols] C-contiguous."""
        left = np.empty((self.rows, self.cols), np.float32)
        right = np.empty((self.rows, self.cols), np.float32)
        for y0, d in zip(self.band_start, self.band_disparity):
            y1 = min(y0 + BAND_ROWS, self.rows)
            off = (k * d) // 4
            left[y0:y1] = self.canvas[y0:y1, off:off + self.cols]
            right[y0:y1] = self.canvas[y0:y1, off + d:off + d + self.cols]
        if noise_seed is None:
            noise_seed = 1000 * (self.seed + 1) + 2 * k
        out = []
        for img, s in ((left, noise_seed), (right, noise_seed + 1)):
            rng = np.random.default_rng(s)
            img = img + rng.standard_normal(img.shape, dtype=np.float32) * np.float32(sigma)
            out.append(np.clip(np.rint(img), 0, 255).astype(np.uint8))
        return out[0], out[1]


def band_world_pair(shape: str | tuple, seed: int, k: int = 0):
    """Independent stereo pair `seed` (config 3 uses seeds 0..4095)."""
    cols, rows = (CAMERAS[shape][:2] if isinstance(shape, str) else shape)
    return BandWorld(cols, rows, seed).pair(k)


def band_world_batch(shape: str | tuple, seeds, workers: int = 0):
    """-> (left[B,rows,cols], right[B,rows,cols]) uint8."""
    seeds = list(seeds)
    cols, rows = (CAMERAS[shape][:2] if isinstance(shape, str) else shape)
    left = np.empty((len(seeds), rows, cols), np.uint8)
    right = np.empty_like(left)
    if workers and len(seeds) > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            for i, (l, r) in enumerate(pool.imap(_pair_job, [((cols, rows), s) for s in seeds], chunksize=4)):
                left[i], right[i] = l, r
    else:
        for i, s in enumerate(seeds):
            left[i], right[i] = band_world_pair((cols, rows), s)
    return left, right


def _pair_job(args):
    return band_world_pair(*args)


def brief_test_table(seed: int = 32) -> np.ndarray:
    """A seeded 256 x 4 (y0, x0, y1, x1) BRIEF test table drawn like the original BRIEF sampling (isotropic Gaussian,
    sigma = patch / 5, clipped to the 48 px patch).  It is NOT opencv_contrib's generated_32.i (absent from this image):
    it exercises the table-parametric BRIEF-32 path; the reference's table is supplied by the host at run time."""
    rng = np.random.default_rng(seed)
    t = np.clip(np.rint(rng.normal(0.0, 48 / 5.0, (256, 4))), -24, 24).astype(np.int8)
    same = (t[:, 0] == t[:, 2]) & (t[:, 1] == t[:, 3])
    t[same, 3] = np.where(t[same, 3] < 24, t[same, 3] + 1, t[same, 3] - 1)
    return t


# ---------------------------------------------------------------------------------------------
# aligner stress set (config 4)
# ---------------------------------------------------------------------------------------------

def _rot(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def true_motion(scale: float = 1.0) -> np.ndarray:
    """T* = rot(0.01,-0.02,0.005 rad) . trans(0.05,-0.02,0.8 m) as a 3x4 [R|t]; `scale` shrinks angles and lengths."""
    R = _rot(0.01 * scale, -0.02 * scale, 0.005 * scale)
    t = R @ (np.array([0.05, -0.02, 0.8]) * scale)
    return np.hstack([R, t[:, None]])


def correspondences(n: int, kind: str = "stereouv", cam: Camera | None = None, seed: int = 424242,
                    outlier_fraction: float = 0.10, max_reliable_depth: float = 15.0):
    """Synthetic 3D-2D correspondences.  Returns a dict of SoA float64 arrays:
    moving[n,3], fixed[n,4|3], omega[n] (StereoUV scalar) or omega_uv[n], omega_d[n] (UVD), wt[n]."""
    cam = cam or camera("kitti")
    rng = np.random.default_rng(seed)
    z = rng.uniform(2.0, 40.0, n)
    u = rng.uniform(0.0, cam.cols, n)
    v = rng.uniform(0.0, cam.rows, n)
    moving = np.stack([(u - cam.cx) / cam.fx * z, (v - cam.cy) / cam.fy * z, z], axis=1)
    T = true_motion()
    p = moving @ T[:, :3].T + T[:, 3]
    abc = p @ cam.K.T
    uvl = abc[:, :2] / abc[:, 2:3]
    abr = abc + cam.baseline
    uvr = abr[:, :2] / abr[:, 2:3]
    has_landmark = rng.random(n) < 0.5
    updates = rng.integers(1, 21, n).astype(np.float64)
    out = {"moving": np.ascontiguousarray(moving), "T_true": T}
    is_out = rng.random(n) < outlier_fraction
    if kind == "stereouv":
        fixed = np.hstack([uvl, uvr]) + rng.normal(0.0, 0.5, (n, 4))
        gross = np.stack([rng.uniform(0, cam.cols, n), rng.uniform(0, cam.rows, n),
                          rng.uniform(0, cam.cols, n), rng.uniform(0, cam.rows, n)], axis=1)
        fixed[is_out] = gross[is_out]
        out["fixed"] = np.ascontiguousarray(fixed)
        out["omega"] = np.where(has_landmark, 1.0 + np.log(updates), 1.0)
        out["wt"] = np.minimum(max_reliable_depth / moving[:, 2], 1.0)
    elif kind == "uvd":
        fixed = np.hstack([uvl + rng.normal(0.0, 0.5, (n, 2)), (p[:, 2] + rng.normal(0.0, 0.02, n) * p[:, 2])[:, None]])
        gross = np.stack([rng.uniform(0, cam.cols, n), rng.uniform(0, cam.rows, n), rng.uniform(2, 40, n)], axis=1)
        fixed[is_out] = gross[is_out]
        unreliable = rng.random(n) < 0.05
        w = np.where(has_landmark, 1.0 + updates, 1.0)
        out["fixed"] = np.ascontiguousarray(fixed)
        out["omega_uv"] = w.copy()
        out["omega_d"] = np.where(unreliable, 0.0, 10.0 * w)
        out["wt"] = np.where(unreliable, 0.0, max_reliable_depth / moving[:, 2])
    else:
        raise ValueError(kind)
    return out


# ---------------------------------------------------------------------------------------------
# landmark histories (Landmark::update)
# ---------------------------------------------------------------------------------------------

LANDMARK_MEASUREMENT = np.dtype([("frame", "<i4"), ("reserved", "<i4"), ("camera_coordinates", "<f8", (3,)),
                                 ("inverse_depth_meters", "<f8")])   # == vslam_landmark_measurement


def landmark_histories(n_landmarks: int, n_frames: int = 40, seed: int = 7, noise: float = 0.02,
                       outlier_fraction: float = 0.05, behind_fraction: float = 0.0):
    """A camera moving forward along z with a small yaw observes `n_landmarks` world points; landmark i has been seen in
    a contiguous run of frames ending at the last one (track lengths 2 .. n_frames, as a track grows in the reference).
    Returns world_to_camera[n_frames, 12], camera_to_world[n_frames, 12] (row-major 3x4), offsets[n+1] (CSR),
    measurements (frame, camera coordinates with N(0, noise * z) error, inverse depth), the initial estimates
    world[n, 3] (the "rude average" of landmark.cpp:21-31), number_of_updates[n] and the true points."""
    rng = np.random.default_rng(seed)
    w2c, c2w = np.zeros((n_frames, 12)), np.zeros((n_frames, 12))
    for f in range(n_frames):
        R = _rot(0.002 * f, 0.01 * f, -0.001 * f)
        t_wc = np.array([0.03 * f, -0.01 * f, 0.7 * f])          # camera position in the world
        c2w[f] = np.hstack([R, t_wc[:, None]]).reshape(12)
        w2c[f] = np.hstack([R.T, (-R.T @ t_wc)[:, None]]).reshape(12)
    last = c2w[-1].reshape(3, 4)
    z = rng.uniform(3.0, 40.0, n_landmarks)
    p_cam = np.stack([rng.uniform(-0.8, 0.8, n_landmarks) * z, rng.uniform(-0.3, 0.3, n_landmarks) * z, z], 1)
    truth = p_cam @ last[:, :3].T + last[:, 3]
    lengths = rng.integers(2, n_frames + 1, n_landmarks)
    offsets = np.zeros(n_landmarks + 1, np.int32)
    offsets[1:] = np.cumsum(lengths)
    ms = np.zeros(int(offsets[-1]), LANDMARK_MEASUREMENT)
    world = np.zeros((n_landmarks, 3))
    for i in range(n_landmarks):
        frames = np.arange(n_frames - lengths[i], n_frames)
        W = w2c[frames].reshape(-1, 3, 4)
        c = np.einsum("fij,j->fi", W[:, :, :3], truth[i]) + W[:, :, 3]
        c = c + rng.normal(0.0, noise, c.shape) * np.abs(c[:, 2:3])
        gross = rng.random(len(frames)) < outlier_fraction
        c[gross] += rng.normal(0.0, 8.0, (int(gross.sum()), 3))
        behind = rng.random(len(frames)) < behind_fraction
        sl = slice(offsets[i], offsets[i + 1])
        ms["frame"][sl] = frames
        ms["camera_coordinates"][sl] = c
        ms["inverse_depth_meters"][sl] = 1.0 / c[:, 2]
        Cw = c2w[frames].reshape(-1, 3, 4)
        world[i] = (np.einsum("fij,fj->fi", Cw[:, :, :3], c) + Cw[:, :, 3]).mean(0)
        if behind.any():                                          # an estimate far behind the first cameras
            world[i] = world[i] - np.array([0.0, 0.0, 200.0])
    n_updates = np.maximum(lengths - rng.integers(1, 4, n_landmarks), 0).astype(np.uint32)
    return {"world_to_camera": w2c, "camera_to_world": c2w, "offsets": offsets, "measurements": ms, "world": world,
            "number_of_updates": n_updates, "truth": truth}
