"""A tracked sequence through vslam_fpg_frame_step, for the ncu launch list (per-kernel durations of one frame graph):
   ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file launches.csv \
       python tools/frame_step_profile.py kitti 12"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from vslam_b200 import api, configs, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "kitti"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
cfg, acfg = configs.BY_NAME[name], configs.ALIGNER_BY_NAME[name]
cam = synth.camera(cfg.camera)
world = synth.BandWorld(cam.cols, cam.rows, 7, max_frames=n)
gen = api.StereoFramePointGenerator(cfg, cam)
gen.frame_step_reset()
T = np.hstack([np.eye(3), np.zeros((3, 1))])
T[0, 3] = -(-cam.bx / cam.fx) / 4
for k in range(n):
    left, right = world.pair(k)
    r = gen.frame_step(left, right, k == 0, T, acfg, False, 15, 25.6, publish_frame_points=True)
    print(k, r["n_previous"], r["n_tracks"], r["n_new_points"], r["aligner_rounds"])
gen.close()
