"""The C-ABI library loads on a CPU-only box and exports every symbol include/vslam_b200.h declares; compute
entry points fail LOUDLY without a GPU (there is no CPU fallback).  Host-side scalar helpers are checked against
the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import tier_a
from vslam_b200 import api, configs, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "vslam_b200.h")).read()
    declared = set(re.findall(r"\b(vslam_[a-z0-9_]+)\s*\(", header))
    assert declared == set(api.EXPORTS), declared ^ set(api.EXPORTS)
    L = api.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.vslam_version()


def test_product_does_not_link_or_import_the_oracle():
    import subprocess
    out = subprocess.run(["ldd", api.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
    undefined = subprocess.run(["nm", "-D", "-C", "--undefined-only", api.LIB_PATH], capture_output=True, text=True).stdout
    assert "vslam" not in undefined, undefined       # every kernel launcher the ABI calls is linked in
    pkg = os.path.join(ROOT, "vslam-pose-estimation-framework_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert not re.search(r"#include\s+[\"<][^\n]*oracle", text), f
                assert "dlopen" not in text and "_build/libvslam_oracle" not in text, f


@pytest.mark.skipif(api.device_count() > 0, reason="checks the no-GPU behaviour")
def test_compute_entry_points_fail_loudly_without_gpu():
    with pytest.raises(api.VslamError) as e:
        api.StereoFramePointGenerator(configs.KITTI, synth.camera("kitti"))
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    with pytest.raises(api.VslamError) as e:
        api.StereoUVAligner(configs.KITTI_ALIGNER, max_points=128)
    assert e.value.code == -2


def test_argument_errors_match_reference_messages():
    cam = synth.camera("kitti")
    c = api.make_config(configs.KITTI, cam)
    c.bx = 10.0                                            # positive b_x -> non-positive baseline
    h = C.c_void_p()
    rc = api.lib().vslam_fpg_create(C.byref(c), 0, C.byref(h))
    assert rc == -1 and b"invalid baseline" in api.lib().vslam_last_error()    # stereo_framepoint_generator.cpp:29-34
    rc = api.lib().vslam_fpg_initialize(None, None, None, 0, 1, None, None)
    assert rc == -1 and b"called with empty frame" in api.lib().vslam_last_error()   # :75-78


def test_host_threshold_controller_matches_oracle():
    f, g = api.lib().vslam_threshold_proposal, tier_a.lib().orc_threshold_proposal
    rng = np.random.default_rng(1)
    for _ in range(500):
        thr = float(rng.integers(5, 101))
        n = int(rng.integers(0, 6000))
        tgt = float(rng.integers(100, 4000))
        tol, chg = float(rng.choice([0.05, 0.1, 0.2])), float(rng.choice([0.1, 0.5, 1.0]))
        assert f(thr, n, tgt, tol, chg, 10.0, 100.0) == g(thr, n, tgt, tol, chg, 10.0, 100.0)


def test_host_solve6_and_v2t_match_oracle():
    rng = np.random.default_rng(2)
    for _ in range(20):
        A = rng.normal(size=(6, 6))
        A = A @ A.T + 0.1 * np.eye(6)
        b = rng.normal(size=6)
        assert np.array_equal(api.solve6(A, b), tier_a.solve6(A, b))
        v = rng.normal(size=6) * 0.1
        assert np.array_equal(api.v2t(v), tier_a.v2t(v))
    assert np.array_equal(api.v2t([0, 0, 0, 3.0, 0, 0]), tier_a.v2t([0, 0, 0, 3.0, 0, 0]))
