#!/bin/bash
# round 2, call AE: full GPU tests, compute-sanitizer (memcheck / racecheck / synccheck / initcheck) on the smallest case of
# every kernel incl. the fused frame, then the bench line
set -x
O=gpurun_out/r2ae
mkdir -p $O
cd /root/repo
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 300 python tools/kernel_zoo.py --small > $O/zoo_small.log 2>&1; echo "rc=$?" >> $O/zoo_small.log
tail -12 $O/zoo_small.log
for tool in memcheck racecheck synccheck initcheck; do
  timeout 1200 compute-sanitizer --tool $tool --log-file $O/sanitizer_$tool.log python tools/kernel_zoo.py --small > $O/zoo_$tool.out 2>&1
  echo "rc=$?" >> $O/zoo_$tool.out
  tail -4 $O/sanitizer_$tool.log
done
python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('$O/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'])
print(d['roofline'].get('kernel_ms_per_step'))
print(json.dumps(d.get('sequence',{}), indent=None)[:3000])
"
