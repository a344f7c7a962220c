#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r21_status.txt
timeout 600 python -m pytest tests/gpu/test_engine_gpu.py tests/gpu/test_zeropp_multi_gpu.py -x -q -k "offload or nvme or pinned" > gpurun_out/r21_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r21_status.txt
timeout 600 python scripts/bench_attn_bias.py > gpurun_out/r21_attn_bias_bench.log 2>&1
echo "attn_bias bench rc=$?" >> gpurun_out/r21_status.txt
timeout 600 python scripts/bench_wq_tc.py > gpurun_out/r21_wq_tc_bench.log 2>&1
echo "wq bench rc=$?" >> gpurun_out/r21_status.txt
tail -3 gpurun_out/r21_tests.log; tail -2 gpurun_out/r21_attn_bias_bench.log | cut -c1-300; cat gpurun_out/r21_status.txt
