#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r4_status.txt
timeout 300 python -m pytest tests/gpu/test_attention_gpu.py -x -q > gpurun_out/r4_attn_tests.log 2>&1
echo "attn tests rc=$?" >> gpurun_out/r4_status.txt
timeout 300 python scripts/bench_attention.py > gpurun_out/r4_attn_bench.log 2>&1
echo "attn bench rc=$?" >> gpurun_out/r4_status.txt
tail -30 gpurun_out/r4_attn_tests.log; tail -28 gpurun_out/r4_attn_bench.log; cat gpurun_out/r4_status.txt
