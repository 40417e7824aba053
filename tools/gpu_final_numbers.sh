#!/bin/bash
# the numbers DESIGN.md / README.md quote for the fused frame at the end of round 2, plus the bench line
O=gpurun_out/r2final4
mkdir -p $O
cd /root/repo
python - <<'PY' > $O/frame_timing.log 2>&1
import json, sys
sys.path.insert(0, ".")
import bench
from vslam_b200 import configs, synth
for name in ("kitti", "euroc", "hd"):
    cfg = configs.BY_NAME[name]
    cam = synth.camera(cfg.camera)
    world = synth.BandWorld(cam.cols, cam.rows, 7, max_frames=40)
    frames = [world.pair(k) for k in range(24)]
    for mode, label in ((0, "stepwise"), (1, "fused"), (2, "fused + prefetch")):
        r = bench.native_sequence(cfg, cam, frames, passes=5, fused=mode)
        print(name, label, json.dumps({k: r[k] for k in r if k in ("frames_per_s", "ms_per_frame", "mean_tracks", "mean_aligner_rounds", "us_initialize_with_feature_download", "us_track", "us_align", "us_compute")} if r else None))
PY
cat $O/frame_timing.log
python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 800 --csv --log-file $O/launches_frame_warm_kitti.csv python tools/frame_step_profile.py kitti 12 > $O/ncu_frame_kitti.log 2>&1
python tools/launch_frame.py $O/launches_frame_warm_kitti.csv > $O/launches_frame_warm_kitti.txt 2>&1
