#!/bin/bash
mkdir -p gpurun_out/probe
cd /root/repo/tools/probe && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o /tmp/lat_probe lat_probe.cu && /tmp/lat_probe | tee /root/repo/gpurun_out/probe/lat_probe.txt
