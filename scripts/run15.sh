#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r15_status.txt
timeout 600 python -m pytest tests/gpu/test_engine_gpu.py -x -q -k "pinned or offload" > gpurun_out/r15_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r15_status.txt
# 70B width, 4 layers, one GPU: the offload pipeline on the native pinned arenas (debug depth)
timeout 600 python bench.py --gpus 1 --model llama3-70b --layers 4 --offload cpu --zero-init --micro-batch 1 --no-exposed --steps 2 --warmup 3 > gpurun_out/r15_l70b_l4.json 2> gpurun_out/r15_l70b_l4.err
echo "l70b-l4 rc=$?" >> gpurun_out/r15_status.txt
tail -5 gpurun_out/r15_tests.log; tail -c 700 gpurun_out/r15_l70b_l4.json; tail -3 gpurun_out/r15_l70b_l4.err | cut -c1-300; cat gpurun_out/r15_status.txt
