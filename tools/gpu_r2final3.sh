#!/bin/bash
O=gpurun_out/r2final2
mkdir -p $O
cd /root/repo
start=$(date +%s)
python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$? in $(( $(date +%s) - start )) s"
python -c "
import json
d=json.load(open('$O/bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
for k,v in d['sequence'].items():
    if isinstance(v,dict) and 'tracked_fused' in v:
        print(k, [ (n, v[n]['frames_per_s'], v[n]['ms_per_frame']) for n in ('tracked_native','tracked_fused','tracked_fused_prefetch') if v.get(n)])
sm=d.get('sequence_multi') or {}
print({k:(sm[k].get('aggregate_frames_per_s') if isinstance(sm[k],dict) else sm[k]) for k in sm if k in ('aggregate_frames_per_s','fused','fused_prefetch')})
"
