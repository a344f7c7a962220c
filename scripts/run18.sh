#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r18_status.txt
timeout 600 python -m pytest tests/gpu/test_attn_bias_gpu.py tests/gpu/test_fpdt_gpu.py -x -q > gpurun_out/r18_attn_bias_tests.log 2>&1
echo "attn_bias tests rc=$?" >> gpurun_out/r18_status.txt
timeout 600 python -m pytest tests/gpu/test_engine_gpu.py -x -q > gpurun_out/r18_engine_tests.log 2>&1
echo "engine tests rc=$?" >> gpurun_out/r18_status.txt
timeout 600 python scripts/bench_attn_bias.py > gpurun_out/r18_attn_bias_bench.log 2>&1
echo "attn_bias bench rc=$?" >> gpurun_out/r18_status.txt
timeout 600 python bench.py --gpus 1 --model llama3-70b --layers 4 --offload cpu --zero-init --micro-batch 1 --no-exposed --steps 2 --warmup 3 > gpurun_out/r18_l70b_l4.json 2> gpurun_out/r18_l70b_l4.err
echo "l70b-l4 rc=$?" >> gpurun_out/r18_status.txt
tail -3 gpurun_out/r18_attn_bias_tests.log; tail -3 gpurun_out/r18_engine_tests.log; grep '^{' gpurun_out/r18_attn_bias_bench.log | cut -c1-420; tail -c 600 gpurun_out/r18_l70b_l4.json; grep host_step gpurun_out/r18_l70b_l4.err | head -2 | cut -c1-300; cat gpurun_out/r18_status.txt
