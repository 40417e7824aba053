"""ctypes binding of the Tier A CPU oracle (oracle/c/vslam_oracle.c).

TEST INFRASTRUCTURE ONLY -- see oracle/c/vslam_oracle.h.  Imported by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs; never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libvslam_oracle.so")

KP = np.dtype([("x", "<f4"), ("y", "<f4"), ("response", "<f4")])
FEATURE = np.dtype([("x", "<f4"), ("y", "<f4"), ("row", "<i4"), ("col", "<i4"), ("index", "<i4"),
                    ("desc", "u1", (32,))])
MATCH = np.dtype([("index_left", "<i4"), ("index_right", "<i4"), ("xl", "<f4"), ("yl", "<f4"), ("xr", "<f4"),
                  ("yr", "<f4"), ("distance", "<i4"), ("epipolar_offset", "<i4"), ("cam", "<f8", (3,))])
TRACKED = np.dtype([("row", "<i4"), ("col", "<i4"), ("has_previous", "<i4"), ("_pad", "<i4"),
                    ("disparity", "<f8"), ("distance", "<f8")])
RECT = np.dtype([("x", "<i4"), ("y", "<i4"), ("w", "<i4"), ("h", "<i4")])
PREVIOUS_POINT = np.dtype([("cam", "<f8", (3,)), ("world", "<f8", (3,)), ("desc_left", "u1", (32,)),
                           ("desc_right", "u1", (32,)), ("epipolar_offset", "<i4"), ("has_landmark", "<i4"),
                           ("keypoint_size", "<f4"), ("_reserved", "<i4")])
TRACK = np.dtype([("index_previous", "<i4"), ("index_left", "<i4"), ("index_right", "<i4"), ("xl", "<f4"),
                  ("yl", "<f4"), ("xr", "<f4"), ("yr", "<f4"), ("distance", "<i4"), ("epipolar_offset", "<i4"),
                  ("projection_left", "<f4", (2,)), ("projection_right", "<f4", (2,)),
                  ("projection_right_corrected", "<f4", (2,)), ("_reserved", "<i4"), ("cam", "<f8", (3,))])
RECOVERED = np.dtype([("index_lost", "<i4"), ("distance", "<i4"), ("xl", "<f4"), ("yl", "<f4"), ("xr", "<f4"),
                      ("yr", "<f4"), ("cam", "<f8", (3,)), ("desc_left", "u1", (32,)), ("desc_right", "u1", (32,))])
assert PREVIOUS_POINT.itemsize == 128 and TRACK.itemsize == 88 and RECOVERED.itemsize == 112


class StereoCamera(C.Structure):
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double), ("bx", C.c_double)]


class AlignerProblem(C.Structure):
    _fields_ = [("n", C.c_int), ("moving", C.c_void_p), ("fixed", C.c_void_p), ("omega", C.c_void_p),
                ("wt", C.c_void_p), ("K", C.c_double * 9), ("baseline", C.c_double * 3), ("rows", C.c_int),
                ("cols", C.c_int), ("min_depth", C.c_double), ("kernel", C.c_double)]


class LinearSystem(C.Structure):
    _fields_ = [("H", C.c_double * 36), ("b", C.c_double * 6), ("total_error", C.c_double), ("inliers", C.c_int),
                ("outliers", C.c_int)]


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, "c", "vslam_oracle.c"), os.path.join(_HERE, "c", "vslam_oracle.h"),
           os.path.join(_HERE, "data", "orb_pattern_31.inc")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_threshold_proposal.restype = C.c_double
        _lib.orc_threshold_proposal.argtypes = [C.c_double, C.c_int] + [C.c_double] * 5
        _lib.orc_triangulation_threshold.restype = C.c_double
        _lib.orc_triangulation_threshold.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double]
        _lib.orc_adjust_thresholds.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p] + [C.c_double] * 5
        _lib.orc_stereo_compute.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_one_round.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_void_p, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_void_p]
        _lib.orc_converge.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_triangulate.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]
        _lib.orc_track.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_recover_points.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double,
                                            C.c_double, C.c_void_p, C.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _img(a):
    a = np.asarray(a)
    assert a.dtype == np.uint8 and a.ndim == 2 and a.strides[1] == 1
    return a


# ------------------------------------------------------------------------------------------------
def detector_regions(rows, cols, nv, nh):
    out = np.zeros(nv * nh, RECT)
    lib().orc_detector_regions(rows, cols, nv, nh, _p(out))
    return out


def bin_grid(rows, cols, bin_size):
    rb, cb = C.c_int(), C.c_int()
    lib().orc_bin_grid(rows, cols, bin_size, C.byref(rb), C.byref(cb))
    return rb.value, cb.value


def fast_detect(img, threshold, cap=None):
    img = _img(img)
    h, w = img.shape
    cap = cap or (w * h // 4 + 16)
    out = np.zeros(cap, KP)
    n = lib().orc_fast_detect(_p(img), img.strides[0], w, h, int(threshold), _p(out), cap)
    assert n <= cap
    return out[:n].copy()


def detect_keypoints(img, nv, nh, thresholds, cap=None):
    """-> (keypoints in reference order, raw count per region)"""
    img = _img(img)
    h, w = img.shape
    cap = cap or (w * h // 4 + 16)
    out = np.zeros(cap, KP)
    counts = np.zeros(nv * nh, np.int32)
    thr = np.ascontiguousarray(thresholds, np.float64)
    n = lib().orc_detect_keypoints(_p(img), img.strides[0], h, w, nv, nh, _p(thr), _p(out), cap, _p(counts))
    assert n <= cap
    return out[:n].copy(), counts


def adjust_thresholds(thresholds, counts_l, counts_r, target, tolerance, max_change, thr_min, thr_max):
    thr = np.ascontiguousarray(thresholds, np.float64).copy()
    cl = np.ascontiguousarray(counts_l, np.int32)
    cr = np.ascontiguousarray(counts_r, np.int32)
    lib().orc_adjust_thresholds(_p(thr), len(thr), _p(cl), _p(cr), float(target), float(tolerance),
                                float(max_change), float(thr_min), float(thr_max))
    return thr


def gauss7_kernel():
    k = np.zeros(7, np.float32)
    lib().orc_gauss7_kernel(_p(k))
    return k


def gauss7_u8(img):
    img = _img(img)
    h, w = img.shape
    out = np.zeros((h, w), np.uint8)
    lib().orc_gauss7_u8(_p(img), img.strides[0], w, h, _p(out), w)
    return out


def orb_compute(img, kps, blurred=None):
    """-> (filtered keypoints, descriptors [n,32])"""
    img = _img(img)
    h, w = img.shape
    kps = np.ascontiguousarray(kps, KP).copy()
    desc = np.zeros((max(len(kps), 1), 32), np.uint8)
    if blurred is not None:
        blurred = _img(blurred)
        n = lib().orc_orb_compute(_p(img), img.strides[0], w, h, _p(blurred), blurred.strides[0], _p(kps), len(kps),
                                  _p(desc))
    else:
        n = lib().orc_orb_compute(_p(img), img.strides[0], w, h, None, 0, _p(kps), len(kps), _p(desc))
    return kps[:n].copy(), desc[:n].copy()


def brief32_compute(img, kps, tests):
    """BriefDescriptorExtractor(32) with a supplied 256 x 4 (y0, x0, y1, x1) test table -> (filtered kps, desc)"""
    img = _img(img)
    h, w = img.shape
    kps = np.ascontiguousarray(kps, KP).copy()
    tests = np.ascontiguousarray(tests, np.int8)
    assert tests.shape == (256, 4) and np.abs(tests.astype(int)).max() <= 24
    desc = np.zeros((max(len(kps), 1), 32), np.uint8)
    n = lib().orc_brief32_compute(_p(img), img.strides[0], w, h, _p(tests), _p(kps), len(kps), _p(desc))
    return kps[:n].copy(), desc[:n].copy()


def hamming256(a, b):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    return lib().orc_hamming256(_p(a), _p(b))


def triangulation_threshold(localizing, n_left, target, max_dist):
    return lib().orc_triangulation_threshold(int(bool(localizing)), int(n_left), int(target), float(max_dist))


def make_features(kps, desc):
    kps = np.ascontiguousarray(kps, KP)
    desc = np.ascontiguousarray(desc, np.uint8)
    out = np.zeros(len(kps), FEATURE)
    if len(kps):
        lib().orc_make_features(_p(kps), _p(desc), len(kps), _p(out))
    return out


def triangulate(cam: StereoCamera, xl, yl, xr, yr):
    out = np.zeros(3)
    lib().orc_triangulate(C.byref(cam), xl, yl, xr, yr, _p(out))
    return out


def stereo_compute(fl, fr, cam: StereoCamera, max_distance, min_disparity, max_epipolar_offset, enable_binning,
                   bin_size, rows, cols, tracked=None):
    """-> dict(matches, winners, remaining_left, remaining_right)"""
    fl = np.ascontiguousarray(fl, FEATURE).copy()
    fr = np.ascontiguousarray(fr, FEATURE).copy()
    nl, nr = C.c_int(len(fl)), C.c_int(len(fr))
    tracked = np.zeros(0, TRACKED) if tracked is None else np.ascontiguousarray(tracked, TRACKED)
    matches = np.zeros(max(len(fl), 1), MATCH)
    winners = np.zeros(max(len(fl), 1) + len(tracked) + 1, np.int32)
    nw = C.c_int()
    n = lib().orc_stereo_compute(_p(fl), C.byref(nl), _p(fr), C.byref(nr), C.byref(cam), float(max_distance),
                                 float(min_disparity), int(max_epipolar_offset), int(bool(enable_binning)),
                                 int(bin_size), int(rows), int(cols), _p(tracked), len(tracked), _p(matches),
                                 _p(winners), C.byref(nw))
    return {"matches": matches[:n].copy(), "winners": winners[:nw.value].copy(),
            "remaining_left": fl[:nl.value].copy(), "remaining_right": fr[:nr.value].copy()}


def track(fl, fr, rows, cols, cam: StereoCamera, previous, T, track_by_appearance, tracking_distance_pixels,
          max_distance_tracking, max_distance_triangulation, min_disparity):
    """StereoFramePointGenerator::track -> dict(tracks, lost, matched_left, matched_right, tracked_landmarks,
    accumulated_distance); matched_* are flags over the positions of fl / fr (what prune() removes)"""
    fl = np.ascontiguousarray(fl, FEATURE)
    fr = np.ascontiguousarray(fr, FEATURE)
    previous = np.ascontiguousarray(previous, PREVIOUS_POINT)
    T = np.ascontiguousarray(T, np.float64).reshape(12)
    n = len(previous)
    tracks = np.zeros(max(n, 1), TRACK)
    lost = np.zeros(max(n, 1), np.int32)
    ml = np.zeros(max(len(fl), 1), np.uint8)
    mr = np.zeros(max(len(fr), 1), np.uint8)
    n_lost, n_lm, acc = C.c_int(), C.c_int(), C.c_double()
    nt = lib().orc_track(_p(fl), len(fl), _p(fr), len(fr), int(rows), int(cols), C.byref(cam), _p(previous), n, _p(T),
                         int(bool(track_by_appearance)), int(tracking_distance_pixels), float(max_distance_tracking),
                         float(max_distance_triangulation), float(min_disparity), _p(tracks), _p(lost),
                         C.byref(n_lost), _p(ml), _p(mr), C.byref(n_lm), C.byref(acc))
    return {"tracks": tracks[:nt].copy(), "lost": lost[:n_lost.value].copy(), "matched_left": ml[:len(fl)].astype(bool),
            "matched_right": mr[:len(fr)].astype(bool), "tracked_landmarks": n_lm.value,
            "accumulated_distance": acc.value}


def recover_points(blurred_left, blurred_right, cam: StereoCamera, lost, world_to_camera_left, min_depth, max_depth,
                   max_distance_tracking, max_distance_triangulation, min_disparity, brief_tests=None):
    """StereoFramePointGenerator::recoverPoints -> RECOVERED records (brief_tests: BRIEF-32 on the RAW frames)"""
    if brief_tests is not None:
        brief_tests = np.ascontiguousarray(brief_tests, np.int8)
    bl, br = _img(blurred_left), _img(blurred_right)
    assert bl.shape == br.shape and bl.strides == br.strides
    rows, cols = bl.shape
    lost = np.ascontiguousarray(lost, PREVIOUS_POINT)
    W = np.ascontiguousarray(world_to_camera_left, np.float64).reshape(12)
    out = np.zeros(max(len(lost), 1), RECOVERED)
    n = lib().orc_recover_points(_p(bl), _p(br), bl.strides[0], rows, cols, C.byref(cam), _p(lost), len(lost), _p(W),
                                 float(min_depth), float(max_depth), float(max_distance_tracking),
                                 float(max_distance_triangulation), float(min_disparity),
                                 None if brief_tests is None else _p(brief_tests), _p(out))
    return out[:n].copy()


# ------------------------------------------------------------------------------------------------
class Aligner:
    """Holds the SoA inputs of one alignment problem (StereoUVAligner / UVDAligner::initialize output)."""

    def __init__(self, kind, moving, fixed, omega, wt, K, baseline, rows, cols, min_depth, kernel):
        self.kind = {"stereouv": 0, "uvd": 1}[kind] if isinstance(kind, str) else int(kind)
        self.moving = np.ascontiguousarray(moving, np.float64)
        self.fixed = np.ascontiguousarray(fixed, np.float64)
        self.omega = np.ascontiguousarray(omega, np.float64)
        self.wt = np.ascontiguousarray(wt, np.float64)
        n = len(self.moving)
        assert self.fixed.shape == (n, 4 if self.kind == 0 else 3)
        assert self.omega.shape == ((n,) if self.kind == 0 else (n, 2))
        p = AlignerProblem()
        p.n = n
        p.moving, p.fixed, p.omega, p.wt = (a.ctypes.data for a in (self.moving, self.fixed, self.omega, self.wt))
        p.K = (C.c_double * 9)(*np.asarray(K, np.float64).ravel())
        p.baseline = (C.c_double * 3)(*np.asarray(baseline, np.float64).ravel())
        p.rows, p.cols, p.min_depth, p.kernel = int(rows), int(cols), float(min_depth), float(kernel)
        self.p = p
        self.errors = np.zeros(n, np.float64)
        self.inliers = np.zeros(n, np.uint8)

    @staticmethod
    def _sys(s):
        return {"H": np.array(s.H).reshape(6, 6), "b": np.array(s.b), "total_error": s.total_error,
                "inliers": s.inliers, "outliers": s.outliers}

    def linearize(self, T, ignore_outliers=False):
        T = np.ascontiguousarray(T, np.float64).reshape(12)
        s = LinearSystem()
        fn = lib().orc_stereouv_linearize if self.kind == 0 else lib().orc_uvd_linearize
        fn(C.byref(self.p), _p(T), int(bool(ignore_outliers)), C.byref(s), _p(self.errors), _p(self.inliers))
        return self._sys(s)

    def one_round(self, T, damping, ignore_outliers=False):
        T = np.ascontiguousarray(T, np.float64).reshape(12).copy()
        s = LinearSystem()
        lib().orc_one_round(self.kind, C.byref(self.p), float(damping), _p(T), int(bool(ignore_outliers)),
                            C.byref(s), _p(self.errors), _p(self.inliers))
        return T.reshape(3, 4), self._sys(s)

    def converge(self, T0, damping, error_delta, max_iterations, min_inliers):
        T = np.ascontiguousarray(T0, np.float64).reshape(12).copy()
        s = LinearSystem()
        info = np.zeros(36)
        rounds = C.c_int()
        ok = lib().orc_converge(self.kind, C.byref(self.p), float(damping), float(error_delta), int(max_iterations),
                                int(min_inliers), _p(T), C.byref(s), _p(self.errors), _p(self.inliers), _p(info),
                                C.byref(rounds))
        out = self._sys(s)
        out.update(T=T.reshape(3, 4), converged=bool(ok), rounds=rounds.value, information=info.reshape(6, 6))
        return out


LANDMARK_MEASUREMENT = np.dtype([("frame", np.int32), ("reserved", np.int32), ("camera_coordinates", np.float64, 3),
                                 ("inverse_depth_meters", np.float64)])


def landmark_update(measurements, world_to_camera, camera_to_world, world, number_of_updates, max_iterations=100,
                    maximum_error_squared_meters=25.0):
    """Landmark::update (landmark.cpp:66-152) -> (world[3], number_of_updates, outcome, iterations)"""
    ms = np.ascontiguousarray(measurements, LANDMARK_MEASUREMENT)
    w2c = np.ascontiguousarray(world_to_camera, np.float64)
    c2w = np.ascontiguousarray(camera_to_world, np.float64)
    x = np.ascontiguousarray(world, np.float64).copy()
    n_up = C.c_uint32(int(number_of_updates))
    it = C.c_int()
    fn = lib().orc_landmark_update
    fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    outcome = fn(_p(ms), len(ms), _p(w2c), _p(c2w), int(max_iterations), float(maximum_error_squared_meters), _p(x),
                 C.byref(n_up), C.byref(it))
    return x, n_up.value, outcome, it.value


def solve3(A, b):
    A = np.ascontiguousarray(A, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    x = np.zeros(3)
    lib().orc_solve3_fullpiv(_p(A), _p(b), _p(x))
    return x


def rotation_to_quaternion(R):
    R = np.ascontiguousarray(R, np.float64)
    q = np.zeros(4)
    lib().orc_rotation_to_quaternion(_p(R), _p(q))
    return q


def format_trajectory(robot_to_world, timestamp=None):
    """one line of writeTrajectoryKITTI (timestamp None) / writeTrajectoryTUM (world_map.cpp:183-252)"""
    T = np.ascontiguousarray(robot_to_world, np.float64).reshape(12)
    buf = C.create_string_buffer(512)
    if timestamp is None:
        n = lib().orc_format_trajectory_kitti(_p(T), buf, 512)
    else:
        fn = lib().orc_format_trajectory_tum
        fn.argtypes = [C.c_double, C.c_void_p, C.c_char_p, C.c_int]
        n = fn(float(timestamp), _p(T), buf, 512)
    return buf.raw[:n].decode()


def solve6(A, b):
    A = np.ascontiguousarray(A, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    x = np.zeros(6)
    lib().orc_solve6_fullpiv(_p(A), _p(b), _p(x))
    return x


def v2t(v):
    v = np.ascontiguousarray(v, np.float64)
    T = np.zeros(12)
    lib().orc_v2t(_p(v), _p(T))
    return T.reshape(3, 4)
