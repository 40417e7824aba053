#!/bin/bash
# 8 GPUs of one box: the bench line (weak scaling; sequence_multi = BASELINE configs[4], stepwise and fused)
mkdir -p gpurun_out/r2n8b
cd /root/repo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2n8b/bench_n8.json 2> gpurun_out/r2n8b/bench_n8.err; echo "rc=$?"
tail -c 600 gpurun_out/r2n8b/bench_n8.err
python -c "
import json
d=json.load(open('gpurun_out/r2n8b/bench_n8.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'])
print(json.dumps(d.get('sequence_multi',{}).get('fused',{}))[:900])
"
