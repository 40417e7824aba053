#!/bin/bash
cd /root/repo
mkdir -p gpurun_out/r2aj
python -m pytest tests/test_gpu_landmark.py tests/test_gpu_cpp_host.py -m gpu -x -q 2>&1 | tail -3
python tools/landmark_timing.py 2>&1 | tee gpurun_out/r2aj/landmark_timing.log
