#!/bin/bash
# round 2, call Y: clocks inside gn_step (timing build on the box only)
set -x
O=gpurun_out/r2y
mkdir -p $O
cd /root/repo
touch vslam-pose-estimation-framework_b200/csrc/aligner.cu
make -C vslam-pose-estimation-framework_b200/csrc EXTRA_aligner="-fmad=false -DVSLAM_GN_TIMING" > $O/make.log 2>&1
timeout 300 python tools/converge_timing.py 2>&1 | grep "gn" | tail -4 > $O/gn_phases.log; cat $O/gn_phases.log
