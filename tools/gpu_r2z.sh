#!/bin/bash
# round 2, call Z: aligner fill fused into track_resolve, prune fused into the cluster Gauss-Newton
set -x
O=gpurun_out/r2z
mkdir -p $O
cd /root/repo
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -15 $O/pytest_gpu.log
timeout 600 python tools/frame_step_timing.py > $O/frame_step_timing.log 2>&1; cut -c1-200 $O/frame_step_timing.log
for shape in kitti hd; do
VSLAM_NO_FRAME_BRANCHES=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 800 --csv --log-file $O/launches_warm_$shape.csv python tools/frame_step_profile.py $shape 12 > $O/ncu_$shape.log 2>&1
done
ls -la $O
