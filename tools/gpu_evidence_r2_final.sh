#!/bin/bash
# round 2, final evidence: (1) GPU tests, (2) the bench line and the reference arm, (3) the ncu launch list of the same
# bench command at 512 pairs, (4) `ncu --set full` of one launch of every kernel of the batched path, (5) the fused
# frame: timing from the C++ runner, warm launch lists (KITTI / 1920x1080), `ncu --set full` of its kernels.
tag=${1:-r2final}
out=gpurun_out/$tag
mkdir -p $out
cd /root/repo
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/smoke.log
python bench.py --steps 5 --warmup 3 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "bench ref rc=$?"
python bench.py --pairs 512 --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $out/bench512.json 2> $out/bench512.err; echo "bench512 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench_pairs512.csv \
  python bench.py --pairs 512 --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -f -o $out/prof_batch \
  -k regex:'fast_nms_kernel|compact_kernel|blur_kernel|describe_tile_kernel|match_kernel|select_strips_kernel|linearize_pairs_kernel|repitch_kernel' \
  --launch-skip 24 -c 16 python bench.py --pairs 512 --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $out/ncu_batch.log 2>&1; echo "ncu batch rc=$?"
python tools/ncu_summary.py raw $out/prof_batch.ncu-rep > $out/prof_batch_summary.txt 2>&1
python tools/ncu_summary.py launches $out/launches_bench_pairs512.csv > $out/launches_summary.txt 2>&1
python tools/ncu_phases.py $out/prof_batch.ncu-rep $((512*1241*376)) > $out/fast_phases.txt 2>&1
timeout 600 python tools/frame_step_timing.py > $out/frame_step_timing.log 2>&1; cut -c1-160 $out/frame_step_timing.log
for shape in kitti hd; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 800 --csv --log-file $out/launches_frame_warm_$shape.csv python tools/frame_step_profile.py $shape 12 > $out/ncu_frame_$shape.log 2>&1
python tools/launch_frame.py $out/launches_frame_warm_$shape.csv > $out/launches_frame_warm_$shape.txt 2>&1
done
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -f -o $out/prof_frame_step \
  -k regex:'compact_frame_kernel|track_search_kernel|track_resolve_kernel|track_emit_kernel|converge_cluster_kernel|select_strips_kernel|frame_assemble_kernel' \
  --launch-skip 40 -c 16 python tools/frame_step_profile.py kitti 12 > $out/ncu_frame_full.log 2>&1; echo "ncu frame rc=$?"
python tools/ncu_summary.py raw $out/prof_frame_step.ncu-rep > $out/prof_frame_step_summary.txt 2>&1
rm -f $out/prof_batch.ncu-rep
ls -la $out
