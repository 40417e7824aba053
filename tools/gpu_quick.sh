#!/bin/bash
# quick check of the fused frame: parity tests of the frame / track / aligner paths, timing, warm launch list
set -x
O=gpurun_out/${1:-quick}
mkdir -p $O
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_frame_step.py tests/test_gpu_track.py tests/test_gpu_aligner.py tests/test_gpu_cpp_host.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
timeout 600 python tools/frame_step_timing.py > $O/frame_step_timing.log 2>&1; cut -c1-200 $O/frame_step_timing.log
for shape in kitti hd; do
VSLAM_NO_FRAME_BRANCHES=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 800 --csv --log-file $O/launches_warm_$shape.csv python tools/frame_step_profile.py $shape 12 > $O/ncu_$shape.log 2>&1
done
# development builds used during round 2 (on the box only; the shipped library has none of them):
#   make -C vslam-pose-estimation-framework_b200/csrc EXTRA_aligner="-fmad=false -DVSLAM_GN_TIMING"    (clock64 phases of the
#     cluster Gauss-Newton and of gn_step, printed by tools/converge_timing.py runs)
#   make -C vslam-pose-estimation-framework_b200/csrc EXTRA_track="-fmad=false -DVSLAM_TRACK_TIMING"   (phases of
#     track_resolve_kernel, printed by tools/frame_step_profile.py runs)
# switches read by the library: VSLAM_NO_FRAME_BRANCHES=1 (one chain of kernels), VSLAM_FRAME_STEP_SYNC=1
# (cudaStreamSynchronize instead of the polled completion word), VSLAM_NO_FRAME_GRAPH=1 (direct launches)
