#!/bin/bash
# round 2, call W: phase clocks of the cluster Gauss-Newton (development build with -DVSLAM_GN_TIMING, on the box only)
set -x
O=gpurun_out/r2w
mkdir -p $O
cd /root/repo
touch vslam-pose-estimation-framework_b200/csrc/aligner.cu
make -C vslam-pose-estimation-framework_b200/csrc EXTRA_aligner="-fmad=false -DVSLAM_GN_TIMING" > $O/make.log 2>&1
timeout 300 python tools/converge_timing.py > $O/converge_timing.log 2>&1; cat $O/converge_timing.log | cut -c1-250
