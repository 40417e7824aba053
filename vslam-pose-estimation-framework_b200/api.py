"""ctypes binding of libvslam_b200.so (include/vslam_b200.h) and a thin Python mirror of the reference's plugin
interface for the hot path, used by the tests and bench.py:

    StereoFramePointGenerator.{configure (ctor), initialize, track, compute, recoverPoints}  <- BaseFramePointGenerator
        (/root/reference/src/framepoint_generation/base_framepoint_generator.h:110-234,
         stereo_framepoint_generator.cpp:16-60,73-133,135-462,464-681,683-869)
    StereoUVAligner / UVDAligner.{initialize, linearize, oneRound, converge, errors, inliers, ...}  <- BaseFrameAligner
        (/root/reference/src/aligners/base_aligner.h:7-71, base_frame_aligner.h:8-41)

Nothing here computes: every method is one C-ABI call into the CUDA library.  There is no CPU fallback; importing
this module without the built library raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvslam_b200.so")

KEYPOINT = np.dtype([("x", "<f4"), ("y", "<f4"), ("response", "<f4")])
FRAMEPOINT = np.dtype([("index_left", "<i4"), ("index_right", "<i4"), ("xl", "<f4"), ("yl", "<f4"), ("xr", "<f4"),
                       ("yr", "<f4"), ("distance", "<i4"), ("epipolar_offset", "<i4"), ("camera", "<f8", (3,))])
TRACKED = np.dtype([("row", "<i4"), ("col", "<i4"), ("has_previous", "<i4"), ("reserved", "<i4"),
                    ("disparity", "<f8"), ("distance", "<f8")])
PREVIOUS_POINT = np.dtype([("camera_left", "<f8", (3,)), ("world", "<f8", (3,)), ("descriptor_left", "u1", (32,)),
                           ("descriptor_right", "u1", (32,)), ("epipolar_offset", "<i4"), ("has_landmark", "<i4"),
                           ("keypoint_size", "<f4"), ("reserved", "<i4")])
LANDMARK_ESTIMATE = np.dtype([("camera", "<f8", (3,)), ("information_scale", "<f8")])
TRACK = np.dtype([("index_previous", "<i4"), ("index_left", "<i4"), ("index_right", "<i4"), ("xl", "<f4"),
                  ("yl", "<f4"), ("xr", "<f4"), ("yr", "<f4"), ("distance", "<i4"), ("epipolar_offset", "<i4"),
                  ("projection_left", "<f4", (2,)), ("projection_right", "<f4", (2,)),
                  ("projection_right_corrected", "<f4", (2,)), ("reserved", "<i4"), ("camera", "<f8", (3,))])
RECOVERED = np.dtype([("index_lost", "<i4"), ("distance", "<i4"), ("xl", "<f4"), ("yl", "<f4"), ("xr", "<f4"),
                      ("yr", "<f4"), ("camera", "<f8", (3,)), ("descriptor_left", "u1", (32,)),
                      ("descriptor_right", "u1", (32,))])
assert FRAMEPOINT.itemsize == 56 and TRACKED.itemsize == 32
assert PREVIOUS_POINT.itemsize == 128 and TRACK.itemsize == 88 and RECOVERED.itemsize == 112
LANDMARK_MEASUREMENT = np.dtype([("frame", "<i4"), ("reserved", "<i4"), ("camera_coordinates", "<f8", (3,)),
                                 ("inverse_depth_meters", "<f8")])
TRACKED_FROM_LAST_TRACK = -1

# every symbol include/vslam_b200.h declares (tests check that the library exports each one)
EXPORTS = """vslam_last_error vslam_version vslam_device_count vslam_host_alloc vslam_host_free
vslam_fpg_create vslam_fpg_destroy vslam_fpg_info vslam_fpg_get_thresholds vslam_fpg_set_thresholds
vslam_fpg_initialize vslam_fpg_get_features vslam_fpg_get_detection_stats vslam_fpg_set_remaining_features
vslam_fpg_reset_features vslam_fpg_graph_launch_count
vslam_fpg_compute vslam_fpg_get_matches vslam_fpg_track vslam_fpg_recover_points vslam_fpg_prune_tracks
vslam_fpg_frame_step vslam_fpg_frame_step_capacity vslam_fpg_frame_step_reset vslam_fpg_frame_step_set_previous
vslam_fpg_frame_step_prefetch vslam_fpg_frame_step_set_landmark_estimates
vslam_fpg_set_profiling vslam_fpg_get_time_consumption vslam_fpg_batch_upload vslam_fpg_batch_run
vslam_fpg_batch_download vslam_fpg_batch_process vslam_fpg_batch_linearize vslam_fpg_batch_get_systems
vslam_fpg_get_kernel_profile vslam_fpg_batch_get_features vslam_fpg_stream vslam_fpg_synchronize
vslam_fpg_launch_count vslam_fpg_debug_keypoint_mask vslam_fpg_debug_blurred vslam_threshold_proposal
vslam_aligner_create vslam_aligner_destroy vslam_aligner_upload vslam_aligner_linearize vslam_aligner_download
vslam_aligner_one_round vslam_aligner_converge vslam_aligner_converge_fused vslam_aligner_linearize_async vslam_aligner_read_system
vslam_aligner_stream vslam_aligner_synchronize vslam_aligner_launch_count vslam_solve6 vslam_v2t
vslam_landmark_optimizer_create vslam_landmark_optimizer_destroy vslam_landmark_optimizer_update
vslam_landmark_map_create vslam_landmark_map_destroy vslam_landmark_map_set_frame_pose vslam_landmark_map_update_frame
vslam_landmark_map_get vslam_landmark_map_size vslam_landmark_map_launch_count
vslam_landmark_optimizer_launch_count vslam_format_trajectory_kitti vslam_format_trajectory_tum vslam_write_trajectory
vslam_solve3""".split()


class FpgConfig(C.Structure):
    _fields_ = [("rows", C.c_int32), ("cols", C.c_int32),
                ("target_number_of_keypoints_tolerance", C.c_double),
                ("detector_threshold_minimum", C.c_int32), ("detector_threshold_maximum", C.c_int32),
                ("detector_threshold_maximum_change", C.c_double),
                ("number_of_detectors_vertical", C.c_int32), ("number_of_detectors_horizontal", C.c_int32),
                ("enable_keypoint_binning", C.c_int32), ("bin_size_pixels", C.c_int32),
                ("maximum_matching_distance_triangulation", C.c_double), ("minimum_disparity_pixels", C.c_double),
                ("maximum_epipolar_search_offset_pixels", C.c_int32),
                ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double), ("bx", C.c_double),
                ("max_keypoints_per_image", C.c_int32), ("max_batch", C.c_int32),
                ("descriptor_type", C.c_int32), ("reserved", C.c_int32), ("brief_tests", C.c_void_p)]


class AlignerParameters(C.Structure):
    _fields_ = [("error_delta_for_convergence", C.c_double), ("maximum_error_kernel", C.c_double),
                ("damping", C.c_double), ("maximum_number_of_iterations", C.c_int32),
                ("minimum_number_of_inliers", C.c_int32)]


class FrameStepParameters(C.Structure):
    _fields_ = [("track_by_appearance", C.c_int32), ("projection_tracking_distance_pixels", C.c_int32),
                ("maximum_descriptor_distance_tracking", C.c_double), ("aligner", AlignerParameters),
                ("enable_inverse_depth_as_information", C.c_int32),
                ("minimum_track_length_for_landmark_creation", C.c_int32),
                ("maximum_reliable_depth_meters", C.c_double), ("minimum_reliable_depth_meters", C.c_double),
                ("publish_frame_points", C.c_int32), ("reserved", C.c_int32)]


class FrameStepResult(C.Structure):
    _fields_ = [("n_left", C.c_int32), ("n_right", C.c_int32), ("n_previous", C.c_int32), ("n_tracked", C.c_int32),
                ("n_lost", C.c_int32), ("n_tracked_landmarks", C.c_int32), ("n_tracks", C.c_int32),
                ("n_new_points", C.c_int32), ("n_matches", C.c_int32), ("aligner_rounds", C.c_int32),
                ("aligner_converged", C.c_int32), ("aligner_inliers", C.c_int32), ("aligner_outliers", C.c_int32),
                ("inliers_only", C.c_int32), ("average_descriptor_distance", C.c_double),
                ("aligner_total_error", C.c_double), ("previous_to_current", C.c_double * 12),
                ("information", C.c_double * 36), ("tracks", C.c_void_p), ("kept", C.c_void_p),
                ("errors", C.c_void_p), ("inliers", C.c_void_p), ("lost", C.c_void_p), ("points", C.c_void_p),
                ("frame_points", C.c_void_p)]


class LinearSystem(C.Structure):
    _fields_ = [("H", C.c_double * 36), ("b", C.c_double * 6), ("total_error", C.c_double),
                ("number_of_inliers", C.c_int32), ("number_of_outliers", C.c_int32)]


class VslamError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("vslam_b200 error %d: %s" % (code, message))
        self.code = code


_lib = None


def lib():
    """The loaded C-ABI library.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C vslam-pose-estimation-framework_b200/csrc); there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.vslam_last_error.restype = C.c_char_p
        L.vslam_version.restype = C.c_char_p
        L.vslam_fpg_stream.restype = C.c_void_p
        L.vslam_aligner_stream.restype = C.c_void_p
        L.vslam_fpg_launch_count.restype = C.c_int64
        L.vslam_aligner_launch_count.restype = C.c_int64
        L.vslam_threshold_proposal.restype = C.c_double
        L.vslam_threshold_proposal.argtypes = [C.c_double, C.c_int32] + [C.c_double] * 5
        vp, i32, sz = C.c_void_p, C.c_int32, C.c_size_t
        L.vslam_host_alloc.argtypes = [vp, sz]
        L.vslam_host_free.argtypes = [vp]
        L.vslam_fpg_create.argtypes = [vp, C.c_int, vp]
        L.vslam_fpg_destroy.argtypes = [vp]
        L.vslam_fpg_info.argtypes = [vp] * 6
        L.vslam_fpg_get_thresholds.argtypes = [vp, vp]
        L.vslam_fpg_set_thresholds.argtypes = [vp, vp]
        L.vslam_fpg_initialize.argtypes = [vp, vp, vp, sz, C.c_int, vp, vp]
        L.vslam_fpg_get_features.argtypes = [vp, C.c_int, vp, vp, i32, vp]
        L.vslam_fpg_get_detection_stats.argtypes = [vp, vp, vp, vp]
        L.vslam_fpg_compute.argtypes = [vp, vp, i32, vp, i32, vp, vp]
        L.vslam_fpg_set_remaining_features.argtypes = [vp, C.c_int, vp, i32]
        L.vslam_fpg_get_matches.argtypes = [vp, vp, i32, vp]
        L.vslam_fpg_track.argtypes = [vp, vp, i32, vp, C.c_int, i32, C.c_double, vp, i32, vp, vp, vp, vp, vp]
        L.vslam_fpg_recover_points.argtypes = [vp, vp, i32, vp, C.c_double, C.c_double, C.c_double, vp, i32, vp]
        L.vslam_fpg_frame_step.argtypes = [vp, vp, vp, sz, C.c_int, vp, vp, vp]
        L.vslam_fpg_frame_step_prefetch.argtypes = [vp, vp, vp, sz]
        L.vslam_fpg_frame_step_set_landmark_estimates.argtypes = [vp, vp, i32]
        L.vslam_fpg_frame_step_capacity.argtypes = [vp]
        L.vslam_fpg_frame_step_capacity.restype = i32
        L.vslam_fpg_frame_step_reset.argtypes = [vp]
        L.vslam_fpg_frame_step_set_previous.argtypes = [vp, vp, i32]
        L.vslam_fpg_set_profiling.argtypes = [vp, C.c_int]
        L.vslam_fpg_get_time_consumption.argtypes = [vp, vp, vp, vp]
        L.vslam_fpg_batch_upload.argtypes = [vp, i32, vp, vp, sz, sz]
        L.vslam_fpg_batch_run.argtypes = [vp, i32, C.c_int]
        L.vslam_fpg_batch_download.argtypes = [vp, i32, vp, i32, vp, vp, vp, vp]
        L.vslam_fpg_batch_process.argtypes = [vp, i32, vp, vp, sz, sz, C.c_int, vp, i32, vp]
        L.vslam_fpg_batch_get_features.argtypes = [vp, i32, C.c_int, vp, vp, i32, vp]
        L.vslam_fpg_batch_linearize.argtypes = [vp, i32, vp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, i32]
        L.vslam_fpg_batch_get_systems.argtypes = [vp, i32, vp, vp, vp, i32]
        L.vslam_fpg_get_kernel_profile.argtypes = [vp, vp, vp]
        L.vslam_fpg_stream.argtypes = [vp]
        L.vslam_fpg_synchronize.argtypes = [vp]
        L.vslam_fpg_launch_count.argtypes = [vp]
        L.vslam_fpg_debug_keypoint_mask.argtypes = [vp, i32, C.c_int, vp]
        L.vslam_fpg_debug_blurred.argtypes = [vp, i32, C.c_int, vp]
        L.vslam_aligner_create.argtypes = [C.c_int, i32, C.c_int, vp]
        L.vslam_aligner_destroy.argtypes = [vp]
        L.vslam_aligner_upload.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, i32, i32, C.c_double]
        L.vslam_aligner_linearize.argtypes = [vp, vp, C.c_int, C.c_double, vp]
        L.vslam_aligner_download.argtypes = [vp, vp, vp]
        L.vslam_aligner_one_round.argtypes = [vp, vp, C.c_int, vp, vp]
        L.vslam_aligner_converge.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.vslam_aligner_converge_fused.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.vslam_aligner_linearize_async.argtypes = [vp, vp, C.c_int, C.c_double]
        L.vslam_aligner_read_system.argtypes = [vp, vp]
        L.vslam_aligner_stream.argtypes = [vp]
        L.vslam_aligner_synchronize.argtypes = [vp]
        L.vslam_aligner_launch_count.argtypes = [vp]
        L.vslam_solve6.argtypes = [vp, vp, vp]
        L.vslam_v2t.argtypes = [vp, vp]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise VslamError(rc, lib().vslam_last_error().decode())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def device_count() -> int:
    return lib().vslam_device_count()


def pinned_empty(shape, dtype=np.uint8) -> np.ndarray:
    """numpy array over page-locked host memory (vslam_host_alloc); freed with the process."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    ptr = C.c_void_p()
    _check(lib().vslam_host_alloc(C.byref(ptr), n))
    buf = (C.c_uint8 * n).from_address(ptr.value)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def make_config(cfg, cam, max_batch=1, max_keypoints=0) -> FpgConfig:
    c = FpgConfig()
    c.rows, c.cols = cam.rows, cam.cols
    c.target_number_of_keypoints_tolerance = cfg.target_number_of_keypoints_tolerance
    c.detector_threshold_minimum = cfg.detector_threshold_minimum
    c.detector_threshold_maximum = cfg.detector_threshold_maximum
    c.detector_threshold_maximum_change = cfg.detector_threshold_maximum_change
    c.number_of_detectors_vertical = cfg.number_of_detectors_vertical
    c.number_of_detectors_horizontal = cfg.number_of_detectors_horizontal
    c.enable_keypoint_binning = int(cfg.enable_keypoint_binning)
    c.bin_size_pixels = cfg.bin_size_pixels
    c.maximum_matching_distance_triangulation = cfg.maximum_matching_distance_triangulation
    c.minimum_disparity_pixels = cfg.minimum_disparity_pixels
    c.maximum_epipolar_search_offset_pixels = cfg.maximum_epipolar_search_offset_pixels
    c.fx, c.fy, c.cx, c.cy, c.bx = cam.fx, cam.fy, cam.cx, cam.cy, cam.bx
    c.max_keypoints_per_image = max_keypoints
    c.max_batch = max_batch
    if getattr(cfg, "descriptor_type", "ORB") == "BRIEF":
        tests = np.ascontiguousarray(cfg.brief_tests, np.int8)
        assert tests.shape == (256, 4)
        c.descriptor_type = 1
        c.brief_tests = tests.ctypes.data
        c._keep = tests                      # the library copies the table inside vslam_fpg_create
    return c


class StereoFramePointGenerator:
    """GPU drop-in for proslam::StereoFramePointGenerator (initialize / compute), single pair and batched."""

    def __init__(self, cfg, cam, device=0, max_batch=1, max_keypoints=0):
        self.cfg, self.cam = cfg, cam
        self._h = C.c_void_p()
        self._c = make_config(cfg, cam, max_batch, max_keypoints)
        _check(lib().vslam_fpg_create(C.byref(self._c), device, C.byref(self._h)))
        n = C.c_int32()
        rb, cb, tg = C.c_int32(), C.c_int32(), C.c_int32()
        reg = np.zeros((64, 4), np.int32)
        _check(lib().vslam_fpg_info(self._h, C.byref(n), _p(reg), C.byref(rb), C.byref(cb), C.byref(tg)))
        self.number_of_detectors = n.value
        self.detector_regions = reg[:n.value].copy()
        self.rows_bin, self.cols_bin, self.target_number_of_keypoints = rb.value, cb.value, tg.value
        self.max_batch = max(1, max_batch)
        self.capacity = min(65535, max(4096, 4 * tg.value) if max_keypoints <= 0 else max_keypoints)
        self.out_capacity = tg.value if cfg.enable_keypoint_binning else self.capacity
        self.number_of_matches = 0

    def close(self):
        if self._h:
            lib().vslam_fpg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- thresholds (FastDetector::getThreshold / setThreshold)
    @property
    def thresholds(self):
        t = np.zeros(self.number_of_detectors)
        _check(lib().vslam_fpg_get_thresholds(self._h, _p(t)))
        return t

    @thresholds.setter
    def thresholds(self, v):
        t = np.ascontiguousarray(v, np.float64)
        assert t.shape == (self.number_of_detectors,)
        _check(lib().vslam_fpg_set_thresholds(self._h, _p(t)))

    # -- StereoFramePointGenerator::initialize(frame, true)
    def initialize(self, left, right, localizing: bool):
        left, right = _image(left, self.cam), _image(right, self.cam)
        assert left.strides[0] == right.strides[0]
        nl, nr = C.c_int32(), C.c_int32()
        _check(lib().vslam_fpg_initialize(self._h, _p(left), _p(right), left.strides[0], int(bool(localizing)),
                                          C.byref(nl), C.byref(nr)))
        return nl.value, nr.value

    def features(self, side: int, pair: int | None = None):
        """(keypoints, descriptors) of frame->keypointsLeft/Right(), descriptorsLeft/Right() in reference order"""
        kps = np.zeros(self.capacity, KEYPOINT)
        desc = np.zeros((self.capacity, 32), np.uint8)
        n = C.c_int32()
        if pair is None:
            _check(lib().vslam_fpg_get_features(self._h, side, _p(kps), _p(desc), self.capacity, C.byref(n)))
        else:
            _check(lib().vslam_fpg_batch_get_features(self._h, pair, side, _p(kps), _p(desc), self.capacity, C.byref(n)))
        return kps[:n.value].copy(), desc[:n.value].copy()

    def detection_stats(self):
        cl = np.zeros(self.number_of_detectors, np.int32)
        cr = np.zeros(self.number_of_detectors, np.int32)
        d = C.c_double()
        _check(lib().vslam_fpg_get_detection_stats(self._h, _p(cl), _p(cr), C.byref(d)))
        return cl, cr, d.value

    @property
    def graph_launch_count(self) -> int:
        lib().vslam_fpg_graph_launch_count.restype = C.c_int64
        return int(lib().vslam_fpg_graph_launch_count(self._h))

    def reset_features(self):
        """initialize(frame, extract_features=False): every feature of the frame takes part again"""
        _check(lib().vslam_fpg_reset_features(self._h))

    def set_remaining_features(self, side, keypoints):
        """the features track() left unmatched in _feature_matcher_left/right (everything else is pruned)"""
        k = np.ascontiguousarray(keypoints, KEYPOINT)
        _check(lib().vslam_fpg_set_remaining_features(self._h, side, _p(k) if len(k) else None, len(k)))

    # -- StereoFramePointGenerator::track(frame, frame_previous, previous_to_current, lost_points, by_appearance)
    def track(self, previous, previous_to_current, track_by_appearance, projection_tracking_distance_pixels,
              maximum_descriptor_distance_tracking):
        """-> dict(tracks, lost, tracked_landmarks, average_descriptor_distance); the matched features are pruned on
        the device, `compute(TRACKED_FROM_LAST_TRACK)` then pre-loads the tracks into the bins without a round trip"""
        previous = np.ascontiguousarray(previous, PREVIOUS_POINT)
        T = np.ascontiguousarray(previous_to_current, np.float64).reshape(12)
        n = len(previous)
        tracks = np.zeros(max(n, 1), TRACK)
        lost = np.zeros(max(n, 1), np.int32)
        nt, nl, nlm, avg = C.c_int32(), C.c_int32(), C.c_int32(), C.c_double()
        _check(lib().vslam_fpg_track(self._h, _p(previous) if n else None, n, _p(T), int(bool(track_by_appearance)),
                                     int(projection_tracking_distance_pixels),
                                     float(maximum_descriptor_distance_tracking), _p(tracks), len(tracks), C.byref(nt),
                                     _p(lost), C.byref(nl), C.byref(nlm), C.byref(avg)))
        self._last_n_tracks = nt.value
        return {"tracks": tracks[:nt.value].copy(), "lost": lost[:nl.value].copy(), "tracked_landmarks": nlm.value,
                "average_descriptor_distance": avg.value}

    # -- StereoFramePointGenerator::recoverPoints(frame, lost_points)
    def recover_points(self, lost, world_to_camera_left, maximum_descriptor_distance_tracking, minimum_depth=0.1,
                       maximum_depth=1000.0):
        lost = np.ascontiguousarray(lost, PREVIOUS_POINT)
        W = np.ascontiguousarray(world_to_camera_left, np.float64).reshape(12)
        out = np.zeros(max(len(lost), 1), RECOVERED)
        n = C.c_int32()
        _check(lib().vslam_fpg_recover_points(self._h, _p(lost) if len(lost) else None, len(lost), _p(W),
                                              float(minimum_depth), float(maximum_depth),
                                              float(maximum_descriptor_distance_tracking), _p(out), len(out),
                                              C.byref(n)))
        return out[:n.value].copy()

    # -- one tracked frame as one device pass (PoseTracker3D::compute's order: initialize, track, aligner, prune, compute)
    def frame_step_capacity(self) -> int:
        return int(lib().vslam_fpg_frame_step_capacity(self._h))

    def frame_step_reset(self):
        _check(lib().vslam_fpg_frame_step_reset(self._h))

    def frame_step_set_previous(self, previous):
        previous = np.ascontiguousarray(previous, PREVIOUS_POINT)
        _check(lib().vslam_fpg_frame_step_set_previous(self._h, _p(previous) if len(previous) else None, len(previous)))

    def frame_step_set_landmark_estimates(self, estimates):
        """one LANDMARK_ESTIMATE per point of the last frame's points(): information_scale != 0 -> the next frame's pose
        optimisation moves `camera` (the landmark estimate in the previous camera frame) with that information scale"""
        estimates = np.ascontiguousarray(estimates, LANDMARK_ESTIMATE)
        _check(lib().vslam_fpg_frame_step_set_landmark_estimates(self._h, _p(estimates) if len(estimates) else None,
                                                                 len(estimates)))

    def frame_step_prefetch(self, left, right):
        """upload of the NEXT frame's images while the current frame runs; consumed by frame_step(None, None, ...)"""
        left, right = _image(left, self.cam), _image(right, self.cam)
        _check(lib().vslam_fpg_frame_step_prefetch(self._h, _p(left), _p(right), left.strides[0]))
        if not hasattr(self, "_staged"):
            self._staged = []
        self._staged.append((left, right))   # the copy is asynchronous: keep the arrays alive until they are consumed

    def frame_step(self, left, right, localizing, previous_to_current_prior, aligner_cfg, track_by_appearance,
                   projection_tracking_distance_pixels, maximum_descriptor_distance_tracking,
                   minimum_track_length_for_landmark_creation=1, publish_frame_points=True):
        """-> dict with the counts, the optimised motion and copies of the result arrays (tracks, kept, errors, inliers,
        lost, points, frame_points)"""
        staged = left is None and right is None   # the oldest pair of frame_step_prefetch()
        if not staged:
            left, right = _image(left, self.cam), _image(right, self.cam)
        T = np.ascontiguousarray(previous_to_current_prior, np.float64).reshape(12)
        p = FrameStepParameters()
        p.track_by_appearance = int(bool(track_by_appearance))
        p.projection_tracking_distance_pixels = int(projection_tracking_distance_pixels)
        p.maximum_descriptor_distance_tracking = float(maximum_descriptor_distance_tracking)
        p.aligner = AlignerParameters(aligner_cfg.error_delta_for_convergence, aligner_cfg.maximum_error_kernel,
                                      aligner_cfg.damping, aligner_cfg.maximum_number_of_iterations,
                                      aligner_cfg.minimum_number_of_inliers)
        p.enable_inverse_depth_as_information = int(aligner_cfg.enable_inverse_depth_as_information)
        p.minimum_track_length_for_landmark_creation = int(minimum_track_length_for_landmark_creation)
        p.maximum_reliable_depth_meters = float(aligner_cfg.maximum_reliable_depth_meters)
        p.minimum_reliable_depth_meters = float(aligner_cfg.minimum_reliable_depth_meters)
        p.publish_frame_points = int(bool(publish_frame_points))
        r = FrameStepResult()
        _check(lib().vslam_fpg_frame_step(self._h, None if staged else _p(left), None if staged else _p(right),
                                          0 if staged else left.strides[0], int(bool(localizing)), _p(T),
                                          C.byref(p), C.byref(r)))
        if staged:
            self._staged.pop(0)

        def view(ptr, n, dtype):
            if n == 0 or not ptr:
                return np.zeros(0, dtype)
            buf = (C.c_uint8 * (n * np.dtype(dtype).itemsize)).from_address(ptr)
            return np.frombuffer(buf, dtype, n).copy()

        out = {f: getattr(r, f) for f, _ in FrameStepResult._fields_[:16]}
        out["previous_to_current"] = np.array(r.previous_to_current).reshape(3, 4)
        out["information"] = np.array(r.information).reshape(6, 6)
        out["tracks"] = view(r.tracks, r.n_tracks, TRACK)
        out["kept"] = view(r.kept, r.n_tracked, np.uint8).astype(bool)
        out["errors"] = view(r.errors, r.n_tracked, np.float64)
        out["inliers"] = view(r.inliers, r.n_tracked, np.uint8).astype(bool)
        out["lost"] = view(r.lost, r.n_lost, np.int32)
        out["points"] = view(r.points, r.n_new_points, FRAMEPOINT)
        out["frame_points"] = view(r.frame_points, r.n_tracks + r.n_new_points, PREVIOUS_POINT)
        return out

    # -- StereoFramePointGenerator::compute(frame)
    def prune_tracks(self, aligner, maximum_error_kernel):
        """PoseTracker3D::_prunePoints for the device-resident tracks of the last track(): -> keep flags [n_tracks] (bool)"""
        n = C.c_int32(0)
        kept = np.zeros(max(self._last_n_tracks, 1), np.uint8)
        fn = lib().vslam_fpg_prune_tracks
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        _check(fn(self._h, aligner._h, float(maximum_error_kernel), C.byref(n), _p(kept)))
        out = kept[:self._last_n_tracks].astype(bool)
        assert int(out.sum()) == n.value
        return out

    def compute(self, tracked=None):
        """tracked: TRACKED records of the points already in frame->points(), or TRACKED_FROM_LAST_TRACK"""
        if isinstance(tracked, int) and tracked == TRACKED_FROM_LAST_TRACK:
            cap = self.out_capacity
            out = np.zeros(cap, FRAMEPOINT)
            n, nm = C.c_int32(), C.c_int32()
            _check(lib().vslam_fpg_compute(self._h, None, TRACKED_FROM_LAST_TRACK, _p(out), cap, C.byref(n),
                                           C.byref(nm)))
            self.number_of_matches = nm.value
            return out[:n.value].copy()
        tracked = np.zeros(0, TRACKED) if tracked is None else np.ascontiguousarray(tracked, TRACKED)
        cap = self.out_capacity + len(tracked)
        out = np.zeros(cap, FRAMEPOINT)
        n, nm = C.c_int32(), C.c_int32()
        _check(lib().vslam_fpg_compute(self._h, _p(tracked) if len(tracked) else None, len(tracked), _p(out), cap,
                                       C.byref(n), C.byref(nm)))
        self.number_of_matches = nm.value
        return out[:n.value].copy()

    def matches(self):
        out = np.zeros(self.capacity, FRAMEPOINT)
        n = C.c_int32()
        _check(lib().vslam_fpg_get_matches(self._h, _p(out), self.capacity, C.byref(n)))
        return out[:n.value].copy()

    # -- chronometers
    def set_profiling(self, on: bool):
        _check(lib().vslam_fpg_set_profiling(self._h, int(on)))

    def time_consumption(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        _check(lib().vslam_fpg_get_time_consumption(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"keypoint_detection": a.value, "descriptor_extraction": b.value, "point_triangulation": c.value}

    # -- batched
    def batch_upload(self, left, right):
        left, right = _batch(left, self.cam), _batch(right, self.cam)
        assert left.shape == right.shape and left.strides == right.strides
        _check(lib().vslam_fpg_batch_upload(self._h, left.shape[0], _p(left), _p(right), left.strides[1], left.strides[0]))
        return left.shape[0]

    def batch_run(self, n_pairs, localizing=True):
        _check(lib().vslam_fpg_batch_run(self._h, n_pairs, int(bool(localizing))))

    def batch_download(self, n_pairs, out=None):
        if out is None:
            out = np.zeros((n_pairs, self.out_capacity), FRAMEPOINT)
        nf, nm = np.zeros(n_pairs, np.int32), np.zeros(n_pairs, np.int32)
        nl, nr = np.zeros(n_pairs, np.int32), np.zeros(n_pairs, np.int32)
        _check(lib().vslam_fpg_batch_download(self._h, n_pairs, _p(out), out.shape[1], _p(nf), _p(nm), _p(nl), _p(nr)))
        return out, nf, nm, nl, nr

    def batch_process(self, left, right, localizing=True, out=None, counts=None):
        """end-to-end call: host images in, framepoints out"""
        left, right = _batch(left, self.cam), _batch(right, self.cam)
        n = left.shape[0]
        if out is None:
            out = np.zeros((n, self.out_capacity), FRAMEPOINT)
        if counts is None:
            counts = np.zeros(n, np.int32)
        _check(lib().vslam_fpg_batch_process(self._h, n, _p(left), _p(right), left.strides[1], left.strides[0],
                                             int(bool(localizing)), _p(out), out.shape[1], _p(counts)))
        return out, counts

    KERNELS = ("fast_nms", "compact", "blur", "describe", "match", "select", "linearize_pairs", "track", "frame_aligner")

    def kernel_profile(self):
        """{kernel: (accumulated device ms, launches)} while profiling was on"""
        ms = np.zeros(len(self.KERNELS))
        n = np.zeros(len(self.KERNELS), np.int64)
        _check(lib().vslam_fpg_get_kernel_profile(self._h, _p(ms), _p(n)))
        return {k: (float(a), int(b)) for k, a, b in zip(self.KERNELS, ms, n)}

    def batch_linearize(self, n_pairs, previous_to_current, aligner_cfg, ignore_outliers=False, rounds=1):
        """StereoUVAligner::initialize + rounds x linearize per pair, over the pair's own new framepoints"""
        T = np.ascontiguousarray(previous_to_current, np.float64).reshape(12)
        _check(lib().vslam_fpg_batch_linearize(self._h, n_pairs, _p(T), int(bool(ignore_outliers)),
                                               float(aligner_cfg.maximum_error_kernel),
                                               float(aligner_cfg.minimum_reliable_depth_meters),
                                               float(aligner_cfg.maximum_reliable_depth_meters),
                                               int(bool(aligner_cfg.enable_inverse_depth_as_information)), int(rounds)))

    def batch_systems(self, n_pairs, with_points=False, raw=False):
        if raw:   # no Python-side conversion: the ctypes array of vslam_linear_system
            if getattr(self, "_sys_buf", None) is None or len(self._sys_buf) != n_pairs:
                self._sys_buf = (LinearSystem * n_pairs)()
            _check(lib().vslam_fpg_batch_get_systems(self._h, n_pairs, C.byref(self._sys_buf), None, None, 0))
            return self._sys_buf
        sys_ = (LinearSystem * n_pairs)()
        errors = inliers = None
        if with_points:
            errors = np.zeros((n_pairs, self.out_capacity))
            inliers = np.zeros((n_pairs, self.out_capacity), np.uint8)
        _check(lib().vslam_fpg_batch_get_systems(self._h, n_pairs, C.byref(sys_), _p(errors), _p(inliers),
                                                 self.out_capacity))
        out = [{"H": np.array(s.H).reshape(6, 6), "b": np.array(s.b), "total_error": s.total_error,
                "inliers": s.number_of_inliers, "outliers": s.number_of_outliers} for s in sys_]
        return (out, errors, inliers) if with_points else out

    def synchronize(self):
        _check(lib().vslam_fpg_synchronize(self._h))

    @property
    def stream(self) -> int:
        return lib().vslam_fpg_stream(self._h) or 0

    @property
    def launch_count(self) -> int:
        return lib().vslam_fpg_launch_count(self._h)

    # -- parity taps
    def debug_keypoint_mask(self, side, pair=0):
        words = np.zeros((self.cam.rows, (self.cam.cols + 31) // 32), np.uint32)
        _check(lib().vslam_fpg_debug_keypoint_mask(self._h, pair, side, _p(words)))
        bits = np.unpackbits(words.view(np.uint8).reshape(self.cam.rows, -1), axis=1, bitorder="little")
        return bits[:, :self.cam.cols].astype(bool)

    def debug_blurred(self, side, pair=0):
        img = np.zeros((self.cam.rows, self.cam.cols), np.uint8)
        _check(lib().vslam_fpg_debug_blurred(self._h, pair, side, _p(img)))
        return img


def make_previous_points(parts, descriptors_left, descriptors_right, has_landmark=1, world=None):
    """frame->points() of a processed frame (TRACK / FRAMEPOINT record arrays in order, e.g. [tracks, new points]) as
    the next frame's track() / recoverPoints() read them: camera coordinates, the two descriptors of each point (rows
    index_left / index_right of descriptorsLeft() / Right()), epipolar offset"""
    n = sum(len(p) for p in parts)
    out = np.zeros(n, PREVIOUS_POINT)
    i = 0
    for p in parts:
        s = slice(i, i + len(p))
        out["camera_left"][s] = p["camera"]
        out["descriptor_left"][s] = descriptors_left[p["index_left"]]
        out["descriptor_right"][s] = descriptors_right[p["index_right"]]
        out["epipolar_offset"][s] = p["epipolar_offset"]
        i += len(p)
    out["world"] = out["camera_left"] if world is None else world
    out["has_landmark"] = has_landmark
    out["keypoint_size"] = 7.0
    return out


def _image(a, cam):
    a = np.asarray(a)
    if a.dtype != np.uint8 or a.shape != (cam.rows, cam.cols) or a.strides[1] != 1:
        raise ValueError("expected a uint8 image of shape (%d, %d)" % (cam.rows, cam.cols))
    return a


def _batch(a, cam):
    a = np.asarray(a)
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[1:] != (cam.rows, cam.cols) or a.strides[2] != 1:
        raise ValueError("expected uint8 images of shape (B, %d, %d)" % (cam.rows, cam.cols))
    return a


class _FrameAligner:
    """BaseFrameAligner mirror: same method names and meaning as the reference (base_aligner.h:26-48)."""
    KIND = 0

    def __init__(self, parameters, max_points=1 << 17, device=0):
        self.parameters = parameters          # configs.AlignerConfig (mutable, like AlignerParameters*)
        self._h = C.c_void_p()
        _check(lib().vslam_aligner_create(self.KIND, max_points, device, C.byref(self._h)))
        self._n = 0
        self._sys = LinearSystem()
        self._T = np.hstack([np.eye(3), np.zeros((3, 1))]).reshape(12).copy()
        self.has_system_converged = False
        self.number_of_rounds = 0
        self.information_matrix = np.eye(6)

    def close(self):
        if self._h:
            lib().vslam_aligner_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _params(self):
        p = AlignerParameters()
        q = self.parameters
        p.error_delta_for_convergence = q.error_delta_for_convergence
        p.maximum_error_kernel = q.maximum_error_kernel
        p.damping = q.damping
        p.maximum_number_of_iterations = q.maximum_number_of_iterations
        p.minimum_number_of_inliers = q.minimum_number_of_inliers
        return p

    def initialize(self, moving, fixed, omega, weights_translation, K, baseline, rows, cols, previous_to_current=None):
        """the buffers StereoUVAligner/UVDAligner::initialize builds from the two frames (a11 in SURVEY 8a)"""
        moving = np.ascontiguousarray(moving, np.float64)
        fixed = np.ascontiguousarray(fixed, np.float64)
        omega = np.ascontiguousarray(omega, np.float64)
        wt = np.ascontiguousarray(weights_translation, np.float64)
        n = len(moving)
        K = np.ascontiguousarray(K, np.float64)
        baseline = np.ascontiguousarray(baseline, np.float64)
        _check(lib().vslam_aligner_upload(self._h, n, _p(moving), _p(fixed), _p(omega), _p(wt), _p(K), _p(baseline),
                                          int(rows), int(cols), float(self.parameters.minimum_reliable_depth_meters)))
        self._n = n
        if previous_to_current is not None:
            self._T = np.ascontiguousarray(previous_to_current, np.float64).reshape(12).copy()

    def _system(self):
        s = self._sys
        return {"H": np.array(s.H).reshape(6, 6), "b": np.array(s.b), "total_error": s.total_error,
                "inliers": s.number_of_inliers, "outliers": s.number_of_outliers}

    def linearize(self, ignore_outliers=False):
        _check(lib().vslam_aligner_linearize(self._h, _p(self._T), int(bool(ignore_outliers)),
                                             float(self.parameters.maximum_error_kernel), C.byref(self._sys)))
        return self._system()

    def oneRound(self, ignore_outliers=False):
        p = self._params()
        _check(lib().vslam_aligner_one_round(self._h, C.byref(p), int(bool(ignore_outliers)), _p(self._T),
                                             C.byref(self._sys)))
        return self._system()

    def converge(self, fused=True):
        """BaseAligner::converge; fused=True runs the whole Gauss-Newton loop as one persistent device kernel,
        fused=False drives linearize round by round from the host like the reference's loop (same results)"""
        p = self._params()
        info = np.zeros(36)
        ok, rounds = C.c_int32(), C.c_int32()
        fn = lib().vslam_aligner_converge_fused if fused else lib().vslam_aligner_converge
        _check(fn(self._h, C.byref(p), _p(self._T), C.byref(self._sys), _p(info), C.byref(ok), C.byref(rounds)))
        self.has_system_converged = bool(ok.value)
        self.number_of_rounds = rounds.value
        if ok.value:
            self.information_matrix = info.reshape(6, 6)
        return self._system()

    def linearize_async(self, ignore_outliers=False):
        _check(lib().vslam_aligner_linearize_async(self._h, _p(self._T), int(bool(ignore_outliers)),
                                                   float(self.parameters.maximum_error_kernel)))

    def read_system(self):
        _check(lib().vslam_aligner_read_system(self._h, C.byref(self._sys)))
        return self._system()

    def synchronize(self):
        _check(lib().vslam_aligner_synchronize(self._h))

    @property
    def stream(self) -> int:
        return lib().vslam_aligner_stream(self._h) or 0

    @property
    def launch_count(self) -> int:
        return lib().vslam_aligner_launch_count(self._h)

    # getters of BaseAligner
    def previousToCurrent(self):
        return self._T.reshape(3, 4).copy()

    def setPreviousToCurrent(self, T):
        self._T = np.ascontiguousarray(T, np.float64).reshape(12).copy()

    def errors(self):
        e = np.zeros(self._n)
        _check(lib().vslam_aligner_download(self._h, _p(e), None))
        return e

    def inliers(self):
        i = np.zeros(self._n, np.uint8)
        _check(lib().vslam_aligner_download(self._h, None, _p(i)))
        return i.astype(bool)

    def numberOfInliers(self):
        return self._sys.number_of_inliers

    def numberOfOutliers(self):
        return self._sys.number_of_outliers

    def numberOfCorrespondences(self):
        return self._n

    def totalError(self):
        return self._sys.total_error

    def averageError(self):
        return self._sys.total_error / self._n


class StereoUVAligner(_FrameAligner):
    KIND = 0


class UVDAligner(_FrameAligner):
    KIND = 1


class LandmarkOptimizer:
    """batched Landmark::update (reference src/types/landmark.cpp:66-152) through the C ABI"""

    def __init__(self, max_landmarks, max_measurements, max_frames, device=0):
        self._h = C.c_void_p()
        _check(lib().vslam_landmark_optimizer_create(int(max_landmarks), int(max_measurements), int(max_frames), device,
                                                     C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            lib().vslam_landmark_optimizer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def update(self, offsets, measurements, world_to_camera, camera_to_world, world, number_of_updates,
               maximum_number_of_iterations=100, maximum_error_squared_meters=25.0):
        """-> (world[n,3], number_of_updates[n], outcome[n], iterations[n]); the inputs are not modified"""
        off = np.ascontiguousarray(offsets, np.int32)
        ms = np.ascontiguousarray(measurements, LANDMARK_MEASUREMENT)
        w2c = np.ascontiguousarray(world_to_camera, np.float64).reshape(-1, 12)
        c2w = np.ascontiguousarray(camera_to_world, np.float64).reshape(-1, 12)
        x = np.ascontiguousarray(world, np.float64).reshape(-1, 3).copy()
        nu = np.ascontiguousarray(number_of_updates, np.uint32).copy()
        n = len(off) - 1
        outcome, iterations = np.zeros(n, np.uint8), np.zeros(n, np.int32)
        fn = lib().vslam_landmark_optimizer_update
        fn.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_uint32,
                       C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _check(fn(self._h, n, _p(off), _p(ms), len(w2c), _p(w2c), _p(c2w), int(maximum_number_of_iterations),
                  float(maximum_error_squared_meters), _p(x), _p(nu), _p(outcome), _p(iterations)))
        return x, nu, outcome, iterations

    @property
    def launch_count(self) -> int:
        lib().vslam_landmark_optimizer_launch_count.restype = C.c_int64
        return int(lib().vslam_landmark_optimizer_launch_count(self._h))


class LandmarkMap:
    """device-resident landmark histories (vslam_landmark_map): PoseTracker3D::_updatePoints of one frame per call"""

    def __init__(self, max_landmarks, max_measurement_blocks, max_frames, device=0):
        self._h = C.c_void_p()
        _check(lib().vslam_landmark_map_create(int(max_landmarks), int(max_measurement_blocks), int(max_frames), device,
                                               C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            lib().vslam_landmark_map_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_frame_pose(self, frame, world_to_camera, camera_to_world):
        w = np.ascontiguousarray(world_to_camera, np.float64).reshape(12)
        c = np.ascontiguousarray(camera_to_world, np.float64).reshape(12)
        _check(lib().vslam_landmark_map_set_frame_pose(self._h, int(frame), _p(w), _p(c)))

    def update_frame(self, frame, world_to_camera, camera_to_world, ids=(), camera_coordinates=(), new_track_offsets=None,
                     new_tracks=None, new_world=None, maximum_number_of_iterations=100, maximum_error_squared_meters=25.0):
        """-> dict(world[n,3], number_of_updates[n], outcome[n], iterations[n], new_ids[n_new])"""
        w = np.ascontiguousarray(world_to_camera, np.float64).reshape(12)
        c = np.ascontiguousarray(camera_to_world, np.float64).reshape(12)
        ids = np.ascontiguousarray(ids, np.int32)
        cam = np.ascontiguousarray(camera_coordinates, np.float64).reshape(-1, 3)
        n = len(ids)
        assert len(cam) == n
        n_new = 0 if new_track_offsets is None else len(new_track_offsets) - 1
        off = np.ascontiguousarray(new_track_offsets if n_new else [0], np.int32)
        trk = np.ascontiguousarray(new_tracks if n_new else np.zeros(0, LANDMARK_MEASUREMENT), LANDMARK_MEASUREMENT)
        nw = np.ascontiguousarray(new_world if n_new else np.zeros((0, 3)), np.float64).reshape(-1, 3)
        world, upd = np.zeros((max(n, 1), 3)), np.zeros(max(n, 1), np.uint32)
        outcome, its = np.zeros(max(n, 1), np.uint8), np.zeros(max(n, 1), np.int32)
        new_ids = np.zeros(max(n_new, 1), np.int32)
        fn = lib().vslam_landmark_map_update_frame
        fn.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                       C.c_void_p, C.c_void_p]
        _check(fn(self._h, int(frame), _p(w), _p(c), n, _p(ids), _p(cam), n_new, _p(off), _p(trk), _p(nw),
                  int(maximum_number_of_iterations), float(maximum_error_squared_meters), _p(world), _p(upd), _p(outcome),
                  _p(its), _p(new_ids)))
        return {"world": world[:n], "number_of_updates": upd[:n], "outcome": outcome[:n], "iterations": its[:n],
                "new_ids": new_ids[:n_new]}

    def get(self, ids):
        ids = np.ascontiguousarray(ids, np.int32)
        n = len(ids)
        world, upd, cnt = np.zeros((max(n, 1), 3)), np.zeros(max(n, 1), np.uint32), np.zeros(max(n, 1), np.int32)
        _check(lib().vslam_landmark_map_get(self._h, n, _p(ids), _p(world), _p(upd), _p(cnt)))
        return world[:n], upd[:n], cnt[:n]

    def __len__(self):
        return int(lib().vslam_landmark_map_size(self._h))

    @property
    def launch_count(self) -> int:
        lib().vslam_landmark_map_launch_count.restype = C.c_int64
        return int(lib().vslam_landmark_map_launch_count(self._h))


def format_trajectory(robot_to_world, timestamp=None) -> str:
    """one line of WorldMap::writeTrajectoryKITTI (timestamp None) / writeTrajectoryTUM"""
    T = np.ascontiguousarray(robot_to_world, np.float64).reshape(12)
    buf = C.create_string_buffer(512)
    if timestamp is None:
        n = lib().vslam_format_trajectory_kitti(_p(T), buf, 512)
    else:
        fn = lib().vslam_format_trajectory_tum
        fn.argtypes = [C.c_double, C.c_void_p, C.c_char_p, C.c_int32]
        n = fn(float(timestamp), _p(T), buf, 512)
    return buf.raw[:n].decode()


def write_trajectory(filename, robot_to_world, timestamps=None):
    T = np.ascontiguousarray(robot_to_world, np.float64).reshape(-1, 12)
    ts = None if timestamps is None else np.ascontiguousarray(timestamps, np.float64)
    _check(lib().vslam_write_trajectory(str(filename).encode(), 0 if ts is None else 1, len(T), _p(T), _p(ts)))


def solve3(A, b):
    A, b, x = np.ascontiguousarray(A, np.float64), np.ascontiguousarray(b, np.float64), np.zeros(3)
    lib().vslam_solve3(_p(A), _p(b), _p(x))
    return x


def solve6(A, b):
    A, b, x = np.ascontiguousarray(A, np.float64), np.ascontiguousarray(b, np.float64), np.zeros(6)
    lib().vslam_solve6(_p(A), _p(b), _p(x))
    return x


def v2t(v):
    v, T = np.ascontiguousarray(v, np.float64), np.zeros(12)
    lib().vslam_v2t(_p(v), _p(T))
    return T.reshape(3, 4)
