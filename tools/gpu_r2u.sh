#!/bin/bash
# round 2, call U: ncu --set full (warm caches) of the serial kernels of the fused frame
set -x
O=gpurun_out/r2u
mkdir -p $O
cd /root/repo
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -f -o $O/prof_frame_step \
  -k regex:'compact_frame_kernel|track_resolve_kernel|select_strips_kernel|converge_cluster_kernel|frame_assemble_kernel|track_search_kernel' \
  --launch-skip 30 -c 12 python tools/frame_step_profile.py kitti 12 > $O/ncu_full.log 2>&1
tail -5 $O/ncu_full.log
ls -la $O
