"""BRIEF-32 (cv::xfeatures2d::BriefDescriptorExtractor::create(32), reference base_framepoint_generator.cpp:186) --
"parity unpinned": opencv_contrib and its generated_32.i test table are not in this image.  What is checked: the C
restatement of xfeatures2d/src/brief.cpp against a literal numpy restatement through cv2.integral (the same OpenCV
primitive the extractor uses), and the parser of generated_32.i on text rendered in that file's format."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import tier_a
from vslam_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("parse_brief_generated", os.path.join(ROOT, "tools", "parse_brief_generated.py"))
pbg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(pbg)


def test_brief32_matches_integral_image_restatement():
    cv2 = pytest.importorskip("cv2")
    left, _ = synth.band_world_pair("euroc", 5)
    tests = synth.brief_test_table(7)
    kps = tier_a.fast_detect(left, 15)
    got_k, got_d = tier_a.brief32_compute(left, kps, tests)
    h, w = left.shape
    keep = (kps["x"] >= 28) & (kps["x"] < w - 28) & (kps["y"] >= 28) & (kps["y"] < h - 28)   # runByImageBorder(28)
    assert np.array_equal(got_k, kps[keep]) and len(got_k) > 1000 and keep.sum() < len(kps)
    s = cv2.integral(left, sdepth=cv2.CV_32S)

    def smoothed(pt_x, pt_y, y, x):   # smoothedSum(), HALF_KERNEL = 4
        iy, ix = int(pt_y + 0.5) + y, int(pt_x + 0.5) + x
        return int(s[iy + 5, ix + 5]) - int(s[iy + 5, ix - 4]) - int(s[iy - 4, ix + 5]) + int(s[iy - 4, ix - 4])

    for i in range(0, len(got_k), 37):
        want = np.zeros(32, np.uint8)
        for t, (y0, x0, y1, x1) in enumerate(tests.tolist()):
            if smoothed(got_k["x"][i], got_k["y"][i], y0, x0) < smoothed(got_k["x"][i], got_k["y"][i], y1, x1):
                want[t // 8] |= 1 << (7 - t % 8)
        assert np.array_equal(got_d[i], want)


def test_generated_32_parser_round_trip():
    table = synth.brief_test_table(1)
    text = pbg.render(table)
    assert "desc[31] = (uchar)(" in text and text.count("SMOOTHED(") == 1 + 512
    assert np.array_equal(pbg.parse(text), table)
    with pytest.raises(ValueError):
        pbg.parse(text.replace("desc[5]", "dsc[5]"))
