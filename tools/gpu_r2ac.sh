#!/bin/bash
# phase clocks of track_resolve_kernel (timing build on the box only)
set -x
O=gpurun_out/r2ac
mkdir -p $O
cd /root/repo
touch vslam-pose-estimation-framework_b200/csrc/track.cu
make -C vslam-pose-estimation-framework_b200/csrc EXTRA_track="-fmad=false -DVSLAM_TRACK_TIMING" > $O/make.log 2>&1
timeout 300 python tools/frame_step_profile.py kitti 8 2>&1 | grep resolve > $O/resolve_kitti.log; cat $O/resolve_kitti.log
timeout 300 python tools/frame_step_profile.py hd 8 2>&1 | grep resolve > $O/resolve_hd.log; cat $O/resolve_hd.log
