#!/bin/bash
# round 2, call S: GPU tests with compact_frame_kernel / one-barrier cluster Gauss-Newton / register 6x6 solve,
# single-sequence timing (stepwise vs fused frame step), launch list of the fused frame
set -x
O=gpurun_out/r2s
mkdir -p $O
cd /root/repo
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 300 python tools/converge_timing.py > $O/converge_timing.log 2>&1; cat $O/converge_timing.log
timeout 600 python tools/frame_step_timing.py > $O/frame_step_timing.log 2>&1; cat $O/frame_step_timing.log
timeout 300 python tools/frame_step_profile.py kitti 12 > $O/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_frame_step_kitti.csv python tools/frame_step_profile.py kitti 12 > $O/ncu.log 2>&1
timeout 300 python tools/frame_step_profile.py hd 12 > $O/plain_hd.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_frame_step_hd.csv python tools/frame_step_profile.py hd 12 > $O/ncu_hd.log 2>&1
ls -la $O
