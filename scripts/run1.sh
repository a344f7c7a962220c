#!/bin/bash
# GPU run 1: new GEMM epilogues (tests), tuning table, step profile, quick bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1_smi.txt 2>&1
timeout 600 python -m pytest tests/gpu/test_gemm_gpu.py -x -q > gpurun_out/r1_gemm_tests.log 2>&1
echo "gemm tests rc=$?" >> gpurun_out/r1_status.txt
timeout 900 python scripts/tune_gemm.py --models llama3-8b --group-m 0,4,8,16 --out gpurun_out/gemm_table.json --report gpurun_out/gemm_tune_report.json > gpurun_out/r1_tune.log 2>&1
echo "tune rc=$?" >> gpurun_out/r1_status.txt
cp gpurun_out/gemm_table.json deepspeed_b200/ops/gemm_table.json 2>/dev/null
timeout 600 python scripts/profile_step.py --out gpurun_out/step_profile_r2a.txt > gpurun_out/r1_profile.log 2>&1
echo "profile rc=$?" >> gpurun_out/r1_status.txt
timeout 600 python bench.py --gpus 1 --steps 6 --warmup 3 > gpurun_out/r1_bench_n1.json 2> gpurun_out/r1_bench_n1.err
echo "bench rc=$?" >> gpurun_out/r1_status.txt
tail -3 gpurun_out/r1_gemm_tests.log; tail -5 gpurun_out/r1_tune.log; cat gpurun_out/r1_bench_n1.json; cat gpurun_out/r1_status.txt
