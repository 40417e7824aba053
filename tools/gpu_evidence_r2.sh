#!/bin/bash
# round 2 evidence: (1) GPU tests, (2) the bench line, (3) the ncu launch list of the same bench command, (4) `ncu --set full`
# of one launch of every kernel of the batched path (512 pairs per step keeps the replay short), (5) the kernel zoo
# (track / recover / aligner / landmark kernels).  Everything lands in gpurun_out/$1.
tag=${1:-r2}
out=gpurun_out/$tag
mkdir -p $out
cd /root/repo
python -c "import __graft_entry__ as g; g.build()" || exit 1
if [ -z "$SKIP_TESTS" ]; then
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
fi
python bench.py --pairs 512 --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $out/bench512.json 2> $out/bench512.err; echo "bench512 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench_pairs512.csv \
  python bench.py --pairs 512 --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -f -o $out/prof_batch \
  -k regex:'fast_nms_kernel|compact_kernel|blur_kernel|describe_tile_kernel|match_kernel|select_strips_kernel|linearize_pairs_kernel|repitch_kernel' \
  --launch-skip 24 -c 16 python bench.py --pairs 512 --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $out/ncu_batch.log 2>&1; echo "ncu batch rc=$?"
timeout 300 python tools/kernel_zoo.py --profile > $out/zoo_profile.log 2>&1; echo "zoo rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -f -o $out/prof_zoo \
  -k regex:'linearize_kernel|converge_kernel|track_search_kernel|track_resolve_kernel|recover_project_kernel|recover_finish_kernel|landmark_update_kernel|describe_kernel|select_strips_kernel|match_kernel|compact_kernel' \
  -c 40 python tools/kernel_zoo.py --profile > $out/ncu_zoo.log 2>&1; echo "ncu zoo rc=$?"
# summaries are made here (the reports together exceed what gpurun copies back); the zoo report is dropped afterwards
python tools/ncu_summary.py raw $out/prof_batch.ncu-rep > $out/prof_batch_summary.txt 2>&1
python tools/ncu_summary.py raw $out/prof_zoo.ncu-rep > $out/prof_zoo_summary.txt 2>&1
python tools/ncu_summary.py launches $out/launches_bench_pairs512.csv > $out/launches_summary.txt 2>&1
python tools/ncu_phases.py $out/prof_batch.ncu-rep $((512*1241*376)) > $out/fast_phases.txt 2>&1
for k in blur_kernel describe_tile_kernel match_kernel select_strips_kernel compact_kernel linearize_pairs_kernel; do
  python tools/ncu_lines.py $out/prof_batch.ncu-rep $k 1.5 > $out/lines_$k.txt 2>&1
done
for k in converge_kernel track_resolve_kernel track_search_kernel landmark_update_kernel linearize_kernel; do
  python tools/ncu_lines.py $out/prof_zoo.ncu-rep $k 1.5 > $out/lines_$k.txt 2>&1
done
rm -f $out/prof_zoo.ncu-rep
ls -la $out
