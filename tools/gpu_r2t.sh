#!/bin/bash
# round 2, call T: frame graph with parallel branches (blur || FAST+compact, match || aligner, tracks' points() || select)
set -x
O=gpurun_out/r2t
mkdir -p $O
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_frame_step.py tests/test_gpu_fpg.py tests/test_gpu_sequence.py tests/test_gpu_cpp_host.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 600 python tools/frame_step_timing.py kitti hd > $O/frame_step_timing.log 2>&1; cut -c1-200 $O/frame_step_timing.log
VSLAM_NO_FRAME_BRANCHES=1 timeout 600 python tools/frame_step_timing.py kitti hd > $O/frame_step_timing_chain.log 2>&1; cut -c1-200 $O/frame_step_timing_chain.log
for shape in kitti hd; do
VSLAM_NO_FRAME_BRANCHES=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 800 --csv --log-file $O/launches_warm_$shape.csv python tools/frame_step_profile.py $shape 12 > $O/ncu_$shape.log 2>&1
done
ls -la $O
