#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/ from OpenCV 4.13 (python cv2).

The reference holds no golden vectors for this path and cannot be built here, so
the fixtures pin the oracle at the boundary the reference delegates to OpenCV
(/root/reference/src/framepoint_generation/base_framepoint_generator.cpp:10-24,195,222,367,434 and
stereo_framepoint_generator.cpp:339):

  fast_orb_<shape>.npz : a synthetic band-world crop, cv2.FastFeatureDetector keypoints (x, y, response)
                         at two thresholds, cv2.ORB_create().compute descriptors, Hamming distances of
                         the first 64 descriptor pairs via cv2.norm(NORM_HAMMING).
Run:  python tools/make_golden.py      (cv2 needed; the tests only need numpy)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import cv2
    from vslam_b200 import synth
    cv2.setNumThreads(0)
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, shape, seed, crop in (("kitti_crop", "kitti", 11, (slice(60, 250), slice(300, 620))),
                                    ("euroc_crop", "euroc", 12, (slice(100, 340), slice(200, 456)))):
        left, _ = synth.band_world_pair(shape, seed)
        img = np.ascontiguousarray(left[crop])
        rec = {"image": img}
        for t in (12, 25):
            kps = cv2.FastFeatureDetector_create(t).detect(img)
            rec["fast%d" % t] = np.array([[k.pt[0], k.pt[1], k.response] for k in kps], np.float32)
        kps = cv2.FastFeatureDetector_create(12).detect(img)
        kps2, desc = cv2.ORB_create().compute(img, kps)
        rec["orb_kps"] = np.array([[k.pt[0], k.pt[1], k.response] for k in kps2], np.float32)
        rec["orb_desc"] = desc
        n = min(64, len(desc) - 1)
        rec["hamming"] = np.array([cv2.norm(desc[i], desc[i + 1], cv2.NORM_HAMMING) for i in range(n)], np.int32)
        rec["cv2_version"] = np.array(cv2.__version__)
        path = os.path.join(out_dir, "fast_orb_%s.npz" % name)
        np.savez_compressed(path, **rec)
        print(path, img.shape, {k: v.shape for k, v in rec.items() if hasattr(v, "shape")}, os.path.getsize(path))


if __name__ == "__main__":
    main()
