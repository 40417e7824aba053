"""Effective parameter sets of the reference's configurations/*.yaml for the hot path.

The drop-in adapters (adapters/) read the reference's own parameter structs directly; this module
holds the same values for the Python harness (tests, bench), i.e. what
`ParameterCollection::parseFromFile` (/root/reference/src/types/parameters.cpp:272-441) leaves in
`StereoFramePointGeneratorParameters` / `AlignerParameters` (/root/reference/src/types/parameters.h:66-95,161-238)
after its parsing quirks (SURVEY.md section 5 "Config / flags"):

* `detector_type` is parsed only in RGB_DEPTH mode (parameters.cpp:341) -> stereo always runs FAST;
* `descriptor_type` BRIEF without opencv_contrib, `BRIEF-256`, `ORB-256` all end in cv::ORB::create()
  (base_framepoint_generator.cpp:187-192,219-224);
* `maximum_matching_distance_triangulation` is parsed as int32_t (parameters.cpp:323): 51.2 fails the
  conversion and the default 0.2*256 = 51.2 stays; integers parse normally.
"""
from __future__ import annotations

import dataclasses


@dataclasses.dataclass
class FramePointGenerationConfig:
    name: str
    camera: str                                   # key into synth.CAMERAS
    # BaseFramePointGeneratorParameters (parameters.h:161-201)
    target_number_of_keypoints_tolerance: float = 0.1
    detector_threshold_minimum: int = 20
    detector_threshold_maximum: int = 100
    detector_threshold_maximum_change: float = 0.1
    number_of_detectors_vertical: int = 1
    number_of_detectors_horizontal: int = 1
    maximum_reliable_depth_meters: float = 15.0
    enable_keypoint_binning: bool = True
    bin_size_pixels: int = 15
    # StereoFramePointGeneratorParameters (parameters.h:204-238)
    maximum_matching_distance_triangulation: float = 0.2 * 256
    minimum_disparity_pixels: float = 1.0
    maximum_epipolar_search_offset_pixels: int = 0
    # descriptor extractor (base_framepoint_generator.cpp:184-224): "ORB" = cv::ORB::create() (what every stereo YAML
    # resolves to without opencv_contrib), "BRIEF" = xfeatures2d::BriefDescriptorExtractor::create(32) (:186), which
    # needs its 256 x 4 (y0, x0, y1, x1) test table (opencv_contrib's generated_32.i; tools/parse_brief_generated.py)
    descriptor_type: str = "ORB"
    brief_tests: object = None


@dataclasses.dataclass
class AlignerConfig:
    # AlignerParameters (parameters.h:66-95)
    error_delta_for_convergence: float = 1e-5
    maximum_error_kernel: float = 10.0
    damping: float = 0.0
    maximum_number_of_iterations: int = 1000
    minimum_number_of_inliers: int = 100
    enable_inverse_depth_as_information: bool = True
    # slam_assembly.cpp:69-70: aligner min/max reliable depth = generator minimum_depth_meters (0.1) /
    # maximum_reliable_depth_meters (BaseAligner defaults 0.01 / 15 at base_aligner.h:62-63 are overridden)
    minimum_reliable_depth_meters: float = 0.1
    maximum_reliable_depth_meters: float = 15.0


# configuration_kitti.yaml:58-134
KITTI = FramePointGenerationConfig(name="kitti", camera="kitti")
KITTI_ALIGNER = AlignerConfig(error_delta_for_convergence=1e-3, maximum_error_kernel=4, damping=5,
                              maximum_number_of_iterations=1000, minimum_number_of_inliers=100)

# configuration_kitti_fast.yaml:23,50,54-59,74,79-81,109-113
KITTI_FAST = FramePointGenerationConfig(name="kitti_fast", camera="kitti", detector_threshold_minimum=15,
                                        detector_threshold_maximum=100, detector_threshold_maximum_change=0.5,
                                        bin_size_pixels=25, maximum_matching_distance_triangulation=60.0)
KITTI_FAST_ALIGNER = AlignerConfig(error_delta_for_convergence=1e-3, maximum_error_kernel=16, damping=0,
                                   maximum_number_of_iterations=1000, minimum_number_of_inliers=0)

# configuration_euroc.yaml:52,56-61,72,78,83,111-112
EUROC = FramePointGenerationConfig(name="euroc", camera="euroc", detector_threshold_minimum=10,
                                   detector_threshold_maximum=30, detector_threshold_maximum_change=1.0,
                                   number_of_detectors_vertical=2, number_of_detectors_horizontal=2,
                                   maximum_reliable_depth_meters=5.0, bin_size_pixels=20,
                                   maximum_matching_distance_triangulation=50.0)
EUROC_ALIGNER = AlignerConfig(error_delta_for_convergence=1e-3, maximum_error_kernel=4, damping=0,
                              maximum_number_of_iterations=1000, minimum_number_of_inliers=100,
                              maximum_reliable_depth_meters=5.0)

# BASELINE.json config 5 (not a reference YAML): 1920x1080, bin 23 -> 84 x 47 bins, thresholds as kitti
HD = FramePointGenerationConfig(name="hd", camera="hd", bin_size_pixels=23)

BY_NAME = {c.name: c for c in (KITTI, KITTI_FAST, EUROC, HD)}
ALIGNER_BY_NAME = {"kitti": KITTI_ALIGNER, "kitti_fast": KITTI_FAST_ALIGNER, "euroc": EUROC_ALIGNER,
                   "hd": KITTI_ALIGNER}
