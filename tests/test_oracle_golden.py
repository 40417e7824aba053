"""Tier A oracle against the committed golden vectors (tests/golden/, made by tools/make_golden.py from
OpenCV 4.13 -- the third-party library the reference delegates FAST / ORB / Hamming to)."""
import os

import numpy as np
import pytest

from oracle import tier_a


@pytest.mark.parametrize("name", ["kitti_crop", "euroc_crop"])
def test_fast_orb_hamming_against_cv2_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "fast_orb_%s.npz" % name))
    img = g["image"]
    for t in (12, 25):
        kp = tier_a.fast_detect(img, t)
        want = g["fast%d" % t]
        got = np.stack([kp["x"], kp["y"], kp["response"]], 1)
        assert got.shape == want.shape
        assert np.array_equal(got, want)          # positions, scores AND row-major order: bit-exact
    kp = tier_a.fast_detect(img, 12)
    kp2, desc = tier_a.orb_compute(img, kp)
    assert np.array_equal(np.stack([kp2["x"], kp2["y"], kp2["response"]], 1), g["orb_kps"])
    assert np.array_equal(desc, g["orb_desc"])    # 256-bit descriptors: bit-exact
    ham = [tier_a.hamming256(desc[i], desc[i + 1]) for i in range(len(g["hamming"]))]
    assert np.array_equal(np.asarray(ham, np.int32), g["hamming"])
