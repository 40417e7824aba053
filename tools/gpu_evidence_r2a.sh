#!/bin/bash
# round 2, call A: GPU tests after the advisor fixes, sanitizer logs on the smallest case of every kernel, and
# `ncu --set full` on the kernels round 1 left without a capture (linearize<0|1>, converge, track, recover, landmark)
set -x
mkdir -p gpurun_out/r2a
cd /root/repo
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2a/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest_gpu.log
tail -3 gpurun_out/r2a/pytest_gpu.log
timeout 300 python tools/kernel_zoo.py --small > gpurun_out/r2a/zoo_small.log 2>&1; echo "rc=$?" >> gpurun_out/r2a/zoo_small.log
cat gpurun_out/r2a/zoo_small.log
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/r2a/sanitizer_$tool.log python tools/kernel_zoo.py --small > gpurun_out/r2a/zoo_$tool.out 2>&1
  echo "rc=$?" >> gpurun_out/r2a/zoo_$tool.out
  tail -5 gpurun_out/r2a/sanitizer_$tool.log
done
timeout 300 python tools/kernel_zoo.py --profile > gpurun_out/r2a/zoo_profile.log 2>&1; echo "rc=$?" >> gpurun_out/r2a/zoo_profile.log
cat gpurun_out/r2a/zoo_profile.log
timeout 1500 ncu --set full --clock-control none --import-source on -f -o gpurun_out/r2a/zoo_full \
  -k regex:'linearize_kernel|converge_kernel|track_search_kernel|track_resolve_kernel|recover_project_kernel|recover_finish_kernel|landmark_update_kernel|describe_kernel|select_strips_kernel|match_kernel|compact_kernel' \
  -c 60 python tools/kernel_zoo.py --profile > gpurun_out/r2a/ncu_zoo.log 2>&1
echo "ncu rc=$?"
ls -la gpurun_out/r2a
