#!/bin/bash
# round 2, call X: warp + shared-memory 6x6 solve: parity, frame timing, phase clocks (timing build on the box only)
set -x
O=gpurun_out/r2x
mkdir -p $O
cd /root/repo
timeout 1200 python -m pytest tests/test_gpu_aligner.py tests/test_gpu_frame_step.py tests/test_gpu_cpp_host.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 600 python tools/frame_step_timing.py kitti hd > $O/frame_step_timing.log 2>&1; cut -c1-200 $O/frame_step_timing.log
timeout 300 python tools/converge_timing.py > $O/converge_timing.log 2>&1; cat $O/converge_timing.log | cut -c1-250
touch vslam-pose-estimation-framework_b200/csrc/aligner.cu
make -C vslam-pose-estimation-framework_b200/csrc EXTRA_aligner="-fmad=false -DVSLAM_GN_TIMING" > $O/make.log 2>&1
timeout 300 python tools/converge_timing.py 2>&1 | grep "gn" | tail -2 > $O/gn_phases.log; cat $O/gn_phases.log
