#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r7_status.txt
timeout 300 python -m pytest tests/gpu/test_attention_gpu.py -x -q > gpurun_out/r7_attn_tests.log 2>&1
echo "attn tests rc=$?" >> gpurun_out/r7_status.txt
timeout 300 python scripts/bench_attention.py > gpurun_out/r7_attn_bench.log 2>&1
echo "attn bench rc=$?" >> gpurun_out/r7_status.txt
timeout 300 python -m pytest tests/gpu/test_misc_ops_gpu.py tests/gpu/test_kernels_gpu.py -x -q > gpurun_out/r7_misc_tests.log 2>&1
echo "misc tests rc=$?" >> gpurun_out/r7_status.txt
tail -4 gpurun_out/r7_attn_tests.log; grep -E "own_|cudnn_" gpurun_out/r7_attn_bench.log; tail -3 gpurun_out/r7_misc_tests.log; cat gpurun_out/r7_status.txt
