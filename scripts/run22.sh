#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r22_status.txt
timeout 900 python -m pytest tests/gpu -x -q -m gpu > gpurun_out/r22_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/r22_status.txt
tail -15 gpurun_out/r22_gpu_tests.log; cat gpurun_out/r22_status.txt
