"""ctypes binding of oracle/_ref/libvslam_ref.so: the reference's own, unmodified hot-path translation units compiled
against the functional stand-ins of oracle/shims (oracle/Makefile, target `_ref`; entry points oracle/ref/ref_harness.h).

TEST INFRASTRUCTURE ONLY.  `available()` is False on a machine that has neither the prebuilt library nor
/root/reference to build it from (tests skip; nothing on the GPU box reads /root/reference).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libvslam_ref.so")
_SO_GPU = os.path.join(_HERE, "_ref", "libvslam_ref_gpu.so")      # + adapters/ + libvslam_b200.so (make _ref_gpu)
REFERENCE_ROOT = "/root/reference"
_lib = None
_lib_gpu = None


class Parameters(C.Structure):
    _fields_ = [
        ("detector_type", C.c_char * 32), ("descriptor_type", C.c_char * 32),
        ("target_number_of_keypoints_tolerance", C.c_double),
        ("detector_threshold_minimum", C.c_double), ("detector_threshold_maximum", C.c_double),
        ("detector_threshold_maximum_change", C.c_double),
        ("number_of_detectors_vertical", C.c_uint32), ("number_of_detectors_horizontal", C.c_uint32),
        ("minimum_projection_tracking_distance_pixels", C.c_int32),
        ("maximum_projection_tracking_distance_pixels", C.c_int32),
        ("minimum_descriptor_distance_tracking", C.c_double), ("maximum_descriptor_distance_tracking", C.c_double),
        ("maximum_reliable_depth_meters", C.c_double), ("maximum_depth_meters", C.c_double),
        ("minimum_depth_meters", C.c_double),
        ("enable_keypoint_binning", C.c_int32), ("bin_size_pixels", C.c_uint32),
        ("maximum_matching_distance_triangulation", C.c_double), ("minimum_disparity_pixels", C.c_double),
        ("maximum_epipolar_search_offset_pixels", C.c_int32), ("use_matches", C.c_int32),
        ("error_delta_for_convergence", C.c_double), ("maximum_error_kernel", C.c_double), ("damping", C.c_double),
        ("maximum_number_of_iterations", C.c_uint32), ("minimum_number_of_inliers", C.c_uint32),
        ("minimum_inlier_ratio", C.c_double), ("enable_inverse_depth_as_information", C.c_int32),
        ("minimum_track_length_for_landmark_creation", C.c_uint32),
        ("minimum_number_of_landmarks_to_track", C.c_uint32),
        ("tunnel_vision_ratio", C.c_double), ("good_tracking_ratio", C.c_double),
        ("maximum_number_of_landmark_recoveries", C.c_uint32), ("enable_landmark_recovery", C.c_int32),
        ("motion_model", C.c_int32),
        ("minimum_delta_angular_for_movement", C.c_double), ("minimum_delta_translational_for_movement", C.c_double),
        ("maximum_error_squared_meters", C.c_double),
    ]


POINT = np.dtype([
    ("xl", "f4"), ("yl", "f4"), ("xr", "f4"), ("yr", "f4"), ("row", "i4"), ("col", "i4"), ("epipolar_offset", "i4"),
    ("index_previous", "i4"), ("disparity", "f8"), ("distance", "f8"), ("cam", "f8", 3), ("robot", "f8", 3),
    ("world", "f8", 3), ("projection_left", "f4", 2), ("projection_right", "f4", 2),
    ("projection_right_corrected", "f4", 2), ("has_landmark", "i4"), ("track_length", "u4"),
    ("landmark_world", "f8", 3), ("landmark_updates", "u4"), ("reserved", "i4"),
    ("desc_left", "u1", 32), ("desc_right", "u1", 32)], align=True)


def build(force: bool = False) -> str | None:
    """make -C oracle _ref (needs /root/reference; a prebuilt library is used as is where the reference is absent)"""
    if not os.path.isdir(REFERENCE_ROOT):
        return _SO if os.path.exists(_SO) else None
    if force:
        subprocess.check_call(["make", "-C", _HERE, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", _HERE, "-j8", "_ref"], stdout=subprocess.DEVNULL)
    return _SO


def build_gpu() -> str | None:
    """make -C oracle _ref_gpu: the reference + adapters/ linked against the product library (must be built first)"""
    if not os.path.isdir(REFERENCE_ROOT):
        return _SO_GPU if os.path.exists(_SO_GPU) else None
    subprocess.check_call(["make", "-C", _HERE, "-j8", "_ref_gpu"], stdout=subprocess.DEVNULL)
    return _SO_GPU


def available() -> bool:
    try:
        return build() is not None
    except Exception:
        return False


def gpu_available() -> bool:
    try:
        return build_gpu() is not None
    except Exception:
        return False


def lib(gpu: bool = False):
    global _lib, _lib_gpu
    if gpu:
        if _lib_gpu is None:
            path = build_gpu()
            if path is None:
                raise RuntimeError("oracle/_ref/libvslam_ref_gpu.so is not built and /root/reference is absent")
            _lib_gpu = _declare(C.CDLL(path))
        return _lib_gpu
    if _lib is None:
        path = build()
        if path is None:
            raise RuntimeError("oracle/_ref is not built and /root/reference is absent")
        _lib = _declare(C.CDLL(path))
    return _lib


def _declare(L):
    if True:
        L.ref_last_error.restype = C.c_char_p
        L.ref_open.restype = C.c_void_p
        L.ref_open.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_double]
        for name in ("ref_fpg_triangulation_distance",):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_void_p]
        for name in ("ref_fpg_seconds", "ref_tracker_seconds"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_void_p, C.c_int]
        L.ref_close.argtypes = [C.c_void_p]
        L.ref_close.restype = None
        assert C.sizeof(Parameters) > 0 and POINT.itemsize == 248, POINT.itemsize
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _t12(T):
    T = np.asarray(T, np.float64)
    if T.shape == (4, 4):
        T = T[:3]
    return np.ascontiguousarray(T.reshape(12))


def effective_values(name):
    """ParameterCollection::parseFromFile(configurations/configuration_<name>.yaml), as ref_get_parameters reports it
    in this repository's container (committed so that the GPU box needs no YAML file)"""
    common = dict(maximum_number_of_iterations=1000, minimum_inlier_ratio=0.0, enable_inverse_depth_as_information=1,
                  maximum_number_of_landmark_recoveries=10, enable_landmark_recovery=1, motion_model=1,
                  minimum_delta_angular_for_movement=0.001, minimum_delta_translational_for_movement=0.01,
                  minimum_number_of_landmarks_to_track=5, target_number_of_keypoints_tolerance=0.1,
                  enable_keypoint_binning=1, minimum_disparity_pixels=1.0, maximum_epipolar_search_offset_pixels=0,
                  maximum_projection_tracking_distance_pixels=50, minimum_depth_meters=0.1, error_delta_for_convergence=1e-3)
    per = {
        "kitti": dict(descriptor_type="BRIEF", detector_threshold_minimum=20, detector_threshold_maximum=100,
                      detector_threshold_maximum_change=0.1, number_of_detectors_vertical=1,
                      number_of_detectors_horizontal=1, minimum_projection_tracking_distance_pixels=15,
                      minimum_descriptor_distance_tracking=25.6, maximum_descriptor_distance_tracking=51.2,
                      maximum_reliable_depth_meters=15.0, maximum_depth_meters=1000.0, bin_size_pixels=15,
                      maximum_matching_distance_triangulation=51.2, maximum_error_kernel=4.0, damping=5.0,
                      minimum_number_of_inliers=100, minimum_track_length_for_landmark_creation=1,
                      tunnel_vision_ratio=0.5, good_tracking_ratio=0.2, maximum_error_squared_meters=100.0),
        "kitti_fast": dict(descriptor_type="BRIEF-256", detector_threshold_minimum=15, detector_threshold_maximum=100,
                           detector_threshold_maximum_change=0.5, number_of_detectors_vertical=1,
                           number_of_detectors_horizontal=1, minimum_projection_tracking_distance_pixels=10,
                           minimum_descriptor_distance_tracking=25.0, maximum_descriptor_distance_tracking=50.0,
                           maximum_reliable_depth_meters=15.0, maximum_depth_meters=1000.0, bin_size_pixels=25,
                           maximum_matching_distance_triangulation=60.0, maximum_error_kernel=16.0, damping=0.0,
                           minimum_number_of_inliers=0, minimum_track_length_for_landmark_creation=2,
                           tunnel_vision_ratio=0.75, good_tracking_ratio=0.2, maximum_error_squared_meters=25.0),
        "euroc": dict(descriptor_type="ORB-256", detector_threshold_minimum=10, detector_threshold_maximum=30,
                      detector_threshold_maximum_change=1.0, number_of_detectors_vertical=2,
                      number_of_detectors_horizontal=2, minimum_projection_tracking_distance_pixels=15,
                      minimum_descriptor_distance_tracking=25.0, maximum_descriptor_distance_tracking=50.0,
                      maximum_reliable_depth_meters=5.0, maximum_depth_meters=100.0, bin_size_pixels=20,
                      maximum_matching_distance_triangulation=50.0, maximum_error_kernel=4.0, damping=0.0,
                      minimum_number_of_inliers=100, minimum_track_length_for_landmark_creation=2,
                      tunnel_vision_ratio=0.5, good_tracking_ratio=0.25, maximum_error_squared_meters=9.0)}
    return {**common, **per[name]}


class Session:
    """One ParameterCollection + cameras + StereoFramePointGenerator + StereoUVAligner / UVDAligner + WorldMap +
    PoseTracker3D of the reference, built as SLAMAssembly builds them."""

    def __init__(self, cam, yaml: str | None = None, gpu: bool = False, **overrides):
        """gpu=True: adapters/ GpuStereoFramePointGenerator + GpuStereoUVAligner take the place of the CPU classes under
        the reference's unmodified tracker (needs a B200 at configure())"""
        self.gpu = gpu
        self.L = lib(gpu)
        self.cam = cam
        K = np.array([[cam.fx, 0, cam.cx], [0, cam.fy, cam.cy], [0, 0, 1]], np.float64)
        self.rows, self.cols = cam.rows, cam.cols
        path = None
        if yaml:
            path = yaml if os.path.isabs(yaml) else os.path.join(REFERENCE_ROOT, "configurations", yaml)
        self.h = self.L.ref_open(path.encode() if path else None, cam.rows, cam.cols, _p(K), C.c_double(cam.bx))
        if not self.h:
            raise RuntimeError(self.L.ref_last_error().decode())
        self.h = C.c_void_p(self.h)
        if overrides:
            p = self.parameters()
            for k, v in overrides.items():
                if not hasattr(p, k):
                    raise AttributeError(k)
                setattr(p, k, v.encode() if isinstance(v, str) else v)
            self._ck(self.L.ref_set_parameters(self.h, C.byref(p)))
        self.configured = False

    def _ck(self, rc):
        if rc < 0:
            raise RuntimeError(self.L.ref_last_error().decode())
        return rc

    def close(self):
        if self.h:
            self.L.ref_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def parameters(self) -> Parameters:
        p = Parameters()
        self._ck(self.L.ref_get_parameters(self.h, C.byref(p)))
        return p

    def configure(self):
        self._ck(self.L.ref_configure_gpu(self.h) if self.gpu else self.L.ref_configure(self.h))
        self.configured = True
        return self

    def reset(self):
        self._ck(self.L.ref_reset(self.h))

    def first_frame(self, left, right, rounds: int, T) -> int:
        """bench.py's per-pair workload on the reference's classes; accumulates self.seconds_pose_optimization"""
        if not hasattr(self, "_pose_seconds"):
            self._pose_seconds = C.c_double(0)
        return self._ck(self.L.ref_first_frame(self.h, _p(left), _p(right), left.strides[0], int(rounds), _p(_t12(T)),
                                               C.byref(self._pose_seconds)))

    @property
    def seconds_pose_optimization(self) -> float:
        return getattr(self, "_pose_seconds", C.c_double(0)).value

    # ---- generator ---------------------------------------------------------------------------------------------
    def initialize(self, left, right, tracking: bool = False):
        left, right = np.ascontiguousarray(left), np.ascontiguousarray(right)
        assert left.shape == (self.rows, self.cols) and left.dtype == np.uint8
        self._ck(self.L.ref_fpg_initialize(self.h, _p(left), _p(right), left.strides[0], int(tracking)))

    def reinitialize(self):
        self._ck(self.L.ref_fpg_reinitialize(self.h))

    def features(self, side: int):
        cap = 200000
        xyr = np.zeros((cap, 3), np.float32)
        desc = np.zeros((cap, 32), np.uint8)
        n = self._ck(self.L.ref_fpg_features(self.h, side, _p(xyr), _p(desc), cap))
        return xyr[:n].copy(), desc[:n].copy()

    def remaining(self, side: int):
        xy = np.zeros((200000, 2), np.float32)
        n = self._ck(self.L.ref_fpg_remaining(self.h, side, _p(xy), len(xy)))
        return xy[:n].copy()

    def thresholds(self):
        out = np.zeros(64, np.float64)
        n = self._ck(self.L.ref_fpg_thresholds(self.h, _p(out), 64))
        return out[:n].copy()

    def triangulation_distance(self) -> float:
        return self.L.ref_fpg_triangulation_distance(self.h)

    def target_number_of_keypoints(self) -> int:
        return self.L.ref_fpg_target_number_of_keypoints(self.h)

    def set_tracking(self, distance_pixels: int, maximum_descriptor_distance: float):
        self._ck(self.L.ref_fpg_set_tracking(self.h, int(distance_pixels), C.c_double(maximum_descriptor_distance)))

    def track(self, previous_to_current, by_appearance: bool):
        lost = np.zeros(200000, np.int32)
        n_lost, n_lm = C.c_int(0), C.c_int(0)
        avg = C.c_double(0)
        n = self._ck(self.L.ref_fpg_track(self.h, _p(_t12(previous_to_current)), int(by_appearance), _p(lost),
                                          C.byref(n_lost), C.byref(n_lm), C.byref(avg)))
        return {"n_tracks": n, "lost": lost[:n_lost.value].copy(), "number_of_tracked_landmarks": n_lm.value,
                "average_descriptor_distance": avg.value}

    def recover(self) -> int:
        return self._ck(self.L.ref_fpg_recover(self.h))

    def compute(self) -> int:
        return self._ck(self.L.ref_fpg_compute(self.h))

    def points(self, previous: bool = False):
        out = np.zeros(200000, POINT)
        n = self._ck(self.L.ref_frame_points(self.h, int(previous), _p(out), len(out)))
        return out[:n].copy()

    def set_pose(self, robot_to_world):
        self._ck(self.L.ref_frame_set_pose(self.h, _p(_t12(robot_to_world))))

    def make_landmarks(self, every_nth: int = 1) -> int:
        return self._ck(self.L.ref_frame_make_landmarks(self.h, every_nth))

    def generator_seconds(self):
        return {k: self.L.ref_fpg_seconds(self.h, i)
                for i, k in enumerate(("keypoint_detection", "descriptor_extraction", "point_triangulation"))}

    # ---- aligners ------------------------------------------------------------------------------------------------
    def aligner_load(self, kind: int, moving, fixed, omega, wt, baseline=(0, 0, 0), min_depth=0.1):
        moving = np.ascontiguousarray(moving, np.float64)
        fixed = np.ascontiguousarray(fixed, np.float64)
        omega = np.ascontiguousarray(omega, np.float64)
        wt = np.ascontiguousarray(wt, np.float64)
        b = np.ascontiguousarray(baseline, np.float64)
        self._n = {**getattr(self, "_n", {}), kind: len(moving)}
        self._ck(self.L.ref_aligner_load(self.h, kind, len(moving), _p(moving), _p(fixed), _p(omega), _p(wt), _p(b),
                                         C.c_double(min_depth)))

    def aligner_set_pose(self, kind: int, T):
        self._ck(self.L.ref_aligner_set_pose(self.h, kind, _p(_t12(T))))

    def aligner_linearize(self, kind: int, ignore_outliers=False):
        self._ck(self.L.ref_aligner_linearize(self.h, kind, int(ignore_outliers)))
        return self.aligner_state(kind)

    def aligner_one_round(self, kind: int, ignore_outliers=False):
        self._ck(self.L.ref_aligner_one_round(self.h, kind, int(ignore_outliers)))
        return self.aligner_state(kind)

    def aligner_converge(self, kind: int):
        converged = self._ck(self.L.ref_aligner_converge(self.h, kind))
        st = self.aligner_state(kind)
        st["converged"] = bool(converged)
        st["rounds"] = self.L.ref_aligner_rounds(self.h, kind)
        return st

    def aligner_state(self, kind: int):
        n = self.L.ref_aligner_count(self.h, kind)
        H, b, T, info = np.zeros(36), np.zeros(6), np.zeros(12), np.zeros(36)
        errors, flags = np.zeros(max(n, 1)), np.zeros(max(n, 1), np.uint8)
        total = C.c_double(0)
        inl, out = C.c_int(0), C.c_int(0)
        self._ck(self.L.ref_aligner_state(self.h, kind, _p(H), _p(b), C.byref(total), C.byref(inl), C.byref(out), _p(T),
                                          _p(errors), _p(flags), _p(info)))
        return {"H": H.reshape(6, 6), "b": b, "total_error": total.value, "inliers": inl.value, "outliers": out.value,
                "T": T.reshape(3, 4), "errors": errors[:n], "inlier_flags": flags[:n].astype(bool),
                "information": info.reshape(6, 6)}

    def aligner_initialize_frames(self, kind: int, T0, enable_inverse_depth_as_information: bool) -> int:
        return self._ck(self.L.ref_aligner_initialize_frames(self.h, kind, _p(_t12(T0)),
                                                             int(enable_inverse_depth_as_information)))

    def aligner_packed(self, kind: int):
        n = self.L.ref_aligner_count(self.h, kind)
        dim = 4 if kind == 0 else 3
        moving, fixed = np.zeros((max(n, 1), 3)), np.zeros((max(n, 1), dim))
        omega = np.zeros(max(n, 1)) if kind == 0 else np.zeros((max(n, 1), 2))
        wt = np.zeros(max(n, 1))
        self._ck(self.L.ref_aligner_packed(self.h, kind, _p(moving), _p(fixed), _p(omega), _p(wt)))
        return moving[:n], fixed[:n], omega[:n], wt[:n]

    # ---- tracker ---------------------------------------------------------------------------------------------------
    def process(self, left, right) -> int:
        left, right = np.ascontiguousarray(left), np.ascontiguousarray(right)
        return self._ck(self.L.ref_tracker_process(self.h, _p(left), _p(right), left.strides[0]))

    def pose(self):
        T = np.zeros(12)
        self._ck(self.L.ref_tracker_pose(self.h, _p(T)))
        return T.reshape(3, 4)

    def status(self) -> int:
        return self.L.ref_tracker_status(self.h)

    def counts(self):
        a, b, c = C.c_int(0), C.c_int(0), C.c_int(0)
        self._ck(self.L.ref_tracker_counts(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"tracked_points": a.value, "landmarks": b.value, "frame_points": c.value}

    def tracker_seconds(self):
        names = ("tracking", "track_creation", "pose_optimization", "landmark_optimization", "point_recovery")
        return {k: self.L.ref_tracker_seconds(self.h, i) for i, k in enumerate(names)}

    def aligner_rounds(self, kind: int = 0) -> int:
        return self.L.ref_aligner_rounds(self.h, kind)

    def write_trajectory(self, fmt: str, filename: str):
        self._ck(self.L.ref_write_trajectory(self.h, 0 if fmt == "kitti" else 1, filename.encode()))

    # ---- landmark ----------------------------------------------------------------------------------------------------
    def landmark_run(self, frame_index, camera_coordinates, robot_to_world):
        fi = np.ascontiguousarray(frame_index, np.int32)
        cc = np.ascontiguousarray(camera_coordinates, np.float64)
        poses = np.ascontiguousarray(np.asarray(robot_to_world, np.float64).reshape(-1, 12))
        world = np.zeros(3)
        nu = C.c_uint32(0)
        self._ck(self.L.ref_landmark_run(self.h, len(fi), _p(fi), _p(cc), len(poses), _p(poses), _p(world), C.byref(nu)))
        return world, nu.value


# ---- OpenCV's own FAST / ORB behind the reference's cv:: calls (timing runs: the reference delegates these to OpenCV) ----
_FAST_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int)
_DESC_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_uint8))
_backend_keepalive = []


def install_cv2_backend(gpu: bool = False):
    """cv::FastFeatureDetector::detect and cv::ORB::compute of the stand-in answer with python cv2 (single-threaded,
    executables/app.cpp:96) instead of tier A: the speed of OpenCV's own SIMD code, the results identical bit for bit
    (tests/test_oracle_vs_cv2.py).  Returns the cv2 version."""
    import cv2
    cv2.setNumThreads(0)
    detectors, orb = {}, cv2.ORB_create()

    def view(ptr, stride, cols, rows):
        buf = (C.c_uint8 * (stride * (rows - 1) + cols)).from_address(ptr)
        return np.lib.stride_tricks.as_strided(np.frombuffer(buf, np.uint8), (rows, cols), (stride, 1))

    def fast(ptr, stride, cols, rows, threshold, xyr, capacity):
        det = detectors.get(threshold)
        if det is None:
            det = detectors[threshold] = cv2.FastFeatureDetector_create(int(threshold))
        kps = det.detect(view(ptr, stride, cols, rows))
        n = len(kps)
        if n > capacity:
            return -1
        if n:
            out = np.ctypeslib.as_array(xyr, (n, 3))
            out[:, :2] = cv2.KeyPoint_convert(kps)
            out[:, 2] = np.fromiter((k.response for k in kps), np.float32, n)
        return n

    def describe(ptr, stride, cols, rows, xyr, n, desc):
        if n == 0:
            return 0
        a = np.ctypeslib.as_array(xyr, (n, 3))
        img = view(ptr, stride, cols, rows)
        kps = cv2.KeyPoint_convert(np.ascontiguousarray(a[:, :2]), size=7.0)
        kps, d = orb.compute(img, kps)
        if d is None:
            return 0
        keep = (a[:, 0] >= 31) & (a[:, 0] < cols - 31) & (a[:, 1] >= 31) & (a[:, 1] < rows - 31)   # ORB's border filter
        kept = a[keep]
        assert len(kept) == len(d)
        a[:len(kept)] = kept
        np.ctypeslib.as_array(desc, (len(d), 32))[:] = d
        return len(d)

    f, o = _FAST_FN(fast), _DESC_FN(describe)
    _backend_keepalive.extend([f, o])
    L = lib(gpu)
    L.vslam_shim_set_backend.argtypes = [_FAST_FN, _DESC_FN, C.c_void_p]
    L.vslam_shim_set_backend(f, o, None)
    return cv2.__version__


def restore_default_backend(gpu: bool = False):
    L = lib(gpu)
    L.vslam_shim_set_backend.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.vslam_shim_set_backend(None, None, None)
