#!/bin/bash
# round 2, last check after the prefetch path: GPU tests, smoke, the bench line and the reference arm
O=gpurun_out/r2final2
mkdir -p $O
cd /root/repo
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
/usr/bin/time -v python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
grep -E "Elapsed|Maximum resident" $O/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench ref rc=$?"
tail -3 $O/pytest_gpu.log
