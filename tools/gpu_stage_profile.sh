#!/bin/bash
# device time per stage of the fused frame (profiling build of the frame graph: one chain with event nodes), and the full
# GPU test suite on the same library
cd /root/repo
mkdir -p gpurun_out/r2stage
python -m pytest tests -m gpu -q 2>&1 | tail -2 | tee gpurun_out/r2stage/pytest_gpu.log
VSLAM_RUNNER_PROFILE=1 python - <<'PY' | tee gpurun_out/r2stage/frame_step_stages.txt
import json, sys
sys.path.insert(0, ".")
import bench
from vslam_b200 import configs, synth
for name in ("kitti", "euroc", "hd"):
    cfg = configs.BY_NAME[name]
    cam = synth.camera(cfg.camera)
    world = synth.BandWorld(cam.cols, cam.rows, 7, max_frames=40)
    frames = [world.pair(k) for k in range(24)]
    r = bench.native_sequence(cfg, cam, frames, passes=5, fused=1)
    print(name, "ms per frame as one chain with events %.4f;" % r["ms_per_frame"], r.get("stage_profile"))
PY
