#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r12_status.txt
timeout 600 python -m pytest tests/gpu/test_attn_bias_gpu.py tests/gpu/test_fpdt_gpu.py tests/gpu/test_misc_ops_gpu.py -x -q > gpurun_out/r12_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r12_status.txt
timeout 600 python scripts/bench_attn_bias.py > gpurun_out/r12_attn_bias_bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/r12_status.txt
tail -25 gpurun_out/r12_tests.log; tail -12 gpurun_out/r12_attn_bias_bench.log | cut -c1-700; cat gpurun_out/r12_status.txt
