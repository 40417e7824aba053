#!/bin/bash
# round 2, call AF: completion word polled in pinned memory instead of cudaStreamSynchronize (A/B)
set -x
O=gpurun_out/r2af
mkdir -p $O
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_frame_step.py tests/test_gpu_cpp_host.py tests/test_gpu_sequence.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
for i in 1 2; do
timeout 600 python tools/frame_step_timing.py kitti euroc hd 2>&1 | grep fused | cut -c1-120 | tee -a $O/poll.log
VSLAM_FRAME_STEP_SYNC=1 timeout 600 python tools/frame_step_timing.py kitti euroc hd 2>&1 | grep fused | cut -c1-120 | tee -a $O/sync.log
done
