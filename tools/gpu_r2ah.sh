#!/bin/bash
# round 2, call AH: vslam_fpg_frame_step_prefetch (next frame uploaded while the current one runs)
set -x
O=gpurun_out/r2ah
mkdir -p $O
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_frame_step.py tests/test_gpu_cpp_host.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -12 $O/pytest_gpu.log
python - <<'PY' > $O/prefetch_timing.log 2>&1
import json, sys
sys.path.insert(0, ".")
import bench
from vslam_b200 import configs, synth
for name in ("kitti", "euroc", "hd"):
    cfg = configs.BY_NAME[name]
    cam = synth.camera(cfg.camera)
    world = synth.BandWorld(cam.cols, cam.rows, 7, max_frames=40)
    frames = [world.pair(k) for k in range(24)]
    for mode in (1, 2):
        r = bench.native_sequence(cfg, cam, frames, passes=5, fused=mode)
        print(name, "fused" if mode == 1 else "fused + prefetch", json.dumps({k: r[k] for k in ("frames_per_s", "ms_per_frame", "mean_tracks", "mean_aligner_rounds", "fused")} if r else None))
PY
cat $O/prefetch_timing.log
