import sys
sys.path.insert(0, ".")
import numpy as np
from vslam_b200 import api, configs, synth
cam = synth.camera("kitti")
for n in (720, 2000):
    c = synth.correspondences(n, "stereouv", cam, seed=3)
    T0 = np.hstack([np.eye(3), np.zeros((3, 1))])
    al = api.StereoUVAligner(configs.KITTI_ALIGNER, max_points=4096)
    al.initialize(c["moving"], c["fixed"], c["omega"], c["wt"], cam.K, cam.baseline, cam.rows, cam.cols, T0)
    for _ in range(3):
        al.setPreviousToCurrent(T0)
        al.converge(fused=True)
    al.close()
