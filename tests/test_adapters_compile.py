"""The adapters (adapters/*.cpp: the C++14 classes that drop libvslam_b200.so into the reference) are
type-checked against the REFERENCE'S OWN HEADERS: every member of Frame / FramePoint / Landmark / Camera /
StereoFramePointGenerator / BaseFrameAligner / PoseTracker3D they touch must exist with a compatible signature.

The reference's third-party dependencies (Eigen, OpenCV C++, srrg_core, srrg_hbst, easy_profiler) are not installed in
this image; tests/stubs/ holds shape-only stand-ins for them (declarations, nothing is linked or run), so this is a
`g++ -std=c++14 -fsyntax-only` check, not a build.  It runs only where /root/reference exists (this container): the GPU
box has no reference tree and skips it."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_SRC = "/root/reference/src"

pytestmark = pytest.mark.skipif(not os.path.isdir(REFERENCE_SRC), reason="the reference tree is not present")


def _type_check(path):
    cmd = ["g++", "-std=c++14", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "tests", "stubs"), "-I", REFERENCE_SRC,
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "adapters"), path]
    return subprocess.run(cmd, capture_output=True, text=True)


@pytest.mark.parametrize("source", ["adapters/gpu_stereo_framepoint_generator.cpp", "adapters/gpu_frame_aligners.cpp",
                                    "tests/cpp/adapter_wiring_check.cpp"])
def test_adapter_sources_type_check_against_the_reference_headers(source):
    r = _type_check(os.path.join(ROOT, source))
    assert r.returncode == 0, r.stderr[-4000:]


@pytest.mark.parametrize("body,needle", [
    ("void f(proslam::Frame* frame_) { frame_->keypointsLeftOfNowhere(); }", "keypointsLeftOfNowhere"),
    ("void f(proslam::PoseTracker3D* t_, proslam::GpuStereoFramePointGenerator* g_) { t_->setAligner(g_); }", "setAligner"),
    ("void f(proslam::GpuStereoUVAligner* a_) { a_->linearize(); }", "linearize"),
])
def test_the_type_check_is_not_vacuous(tmp_path, body, needle):
    """a misuse of the reference's interfaces must be rejected: the stand-in headers do not swallow everything"""
    src = tmp_path / "misuse.cpp"
    src.write_text('#include "position_tracking/pose_tracker_3d.h"\n#include "gpu_frame_aligners.h"\n'
                   '#include "gpu_stereo_framepoint_generator.h"\n' + body + "\n")
    r = _type_check(str(src))
    assert r.returncode != 0 and needle in r.stderr
