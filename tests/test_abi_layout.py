"""The Python harness mirrors every public struct of include/vslam_b200.h (ctypes structures for the parameter / result
blocks, numpy dtypes for the record arrays).  A C program compiled against the header prints sizeof of each struct and
the offset of its last member; both must agree with the mirrors -- on a box without a GPU too, where nothing else would
notice a member added on one side only."""
import ctypes as C
import os
import subprocess

import numpy as np

from vslam_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# C struct, its last member, the mirror
STRUCTS = [
    ("vslam_linear_system", "number_of_outliers", api.LinearSystem),
    ("vslam_fpg_config", None, api.FpgConfig),
    ("vslam_keypoint", "response", api.KEYPOINT),
    ("vslam_framepoint", "camera", api.FRAMEPOINT),
    ("vslam_tracked_point", "distance", api.TRACKED),
    ("vslam_previous_point", "reserved", api.PREVIOUS_POINT),
    ("vslam_track", "camera", api.TRACK),
    ("vslam_recovered_point", None, api.RECOVERED),
    ("vslam_aligner_parameters", "minimum_number_of_inliers", api.AlignerParameters),
    ("vslam_frame_step_parameters", "reserved", api.FrameStepParameters),
    ("vslam_frame_step_result", "frame_points", api.FrameStepResult),
    ("vslam_landmark_estimate", "information_scale", api.LANDMARK_ESTIMATE),
    ("vslam_landmark_measurement", None, api.LANDMARK_MEASUREMENT),
]


def _mirror_size_and_last_offset(mirror, last):
    if isinstance(mirror, np.dtype):
        return mirror.itemsize, (mirror.fields[last][1] if last else None)
    return C.sizeof(mirror), (getattr(mirror, last).offset if last else None)


def test_struct_layouts_match_the_header(tmp_path):
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "vslam_b200.h"', 'int main(void) {']
    for name, last, _ in STRUCTS:
        off = "offsetof(%s, %s)" % (name, last) if last else "(size_t)0"
        lines.append('  printf("%s %%zu %%zu\\n", sizeof(%s), %s);' % (name, name, off))
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split("\n")
    seen = {l.split()[0]: (int(l.split()[1]), int(l.split()[2])) for l in out if l.strip()}
    for name, last, mirror in STRUCTS:
        size, offset = _mirror_size_and_last_offset(mirror, last)
        assert seen[name][0] == size, (name, seen[name][0], size)
        if last:
            assert seen[name][1] == offset, (name, last, seen[name][1], offset)



def test_cpp_hosts_compile_and_link_without_a_gpu(tmp_path):
    """the C++14 layer over the C ABI (include/vslam_b200.hpp) and its two hosts -- the consumer check and the sequence
    runner bench.py times -- build and link against the in-tree library on a CPU-only box (they run in the GPU tests)"""
    pkg = os.path.join(ROOT, "vslam-pose-estimation-framework_b200")
    for source in ("tests/cpp/host_api_check.cpp", "tools/sequence_runner.cpp"):
        exe = tmp_path / os.path.basename(source).replace(".cpp", "")
        subprocess.check_call(["g++", "-std=c++14", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"),
                               os.path.join(ROOT, source), "-o", str(exe), "-L", pkg, "-lvslam_b200", "-Wl,-rpath," + pkg])
        assert exe.exists()
