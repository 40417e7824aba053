#!/bin/bash
# A/B runs of the resident throughput with library switches (environment variables read at vslam_fpg_create)
# usage: tools/ab_bench.sh <out-dir> "VAR=val VAR2=val" ["..." ...]
out=$1; shift
mkdir -p $out
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $out/ab_$i.json 2> $out/ab_$i.err
  python - "$cfg" $out/ab_$i.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); print("%-60s value %.0f ms/step %.3f e2e %.0f"%(sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
