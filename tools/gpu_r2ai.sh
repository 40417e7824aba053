#!/bin/bash
# how much of a fused frame is the host's copy of points() (publish_frame_points)
cd /root/repo
mkdir -p gpurun_out/r2ai
for v in "" "VSLAM_RUNNER_NO_FRAME_POINTS=1"; do
  echo "== $v"
  env $v python tools/frame_step_timing.py kitti hd 2>&1 | grep fused | cut -c1-90
done | tee gpurun_out/r2ai/frame_points.log
