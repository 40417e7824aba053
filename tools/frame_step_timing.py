"""Single tracked sequence, stepwise calls vs vslam_fpg_frame_step, from the C++14 runner (tools/sequence_runner.cpp).
   python tools/frame_step_timing.py [kitti euroc hd]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from vslam_b200 import configs, synth  # noqa: E402

for name in (sys.argv[1:] or ["kitti", "euroc", "hd"]):
    cfg = configs.BY_NAME[name]
    cam = synth.camera(cfg.camera)
    world = synth.BandWorld(cam.cols, cam.rows, 7, max_frames=40)
    frames = [world.pair(k) for k in range(24)]
    for fused in (False, True):
        r = bench.native_sequence(cfg, cam, frames, passes=5, fused=fused)
        print(name, "fused" if fused else "stepwise", json.dumps(r))
