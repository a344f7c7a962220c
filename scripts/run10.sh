#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r10_status.txt
timeout 600 python -m pytest tests/gpu/test_gemm_gpu.py tests/gpu/test_engine_gpu.py -x -q -m gpu > gpurun_out/r10_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/r10_status.txt
timeout 900 python scripts/tune_gemm.py --out gpurun_out/gemm_table_r10.json --report gpurun_out/gemm_tune_report_r10.json > gpurun_out/r10_tune.log 2>&1
echo "tune rc=$?" >> gpurun_out/r10_status.txt
cp gpurun_out/gemm_table_r10.json deepspeed_b200/ops/gemm_table.json
timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 > gpurun_out/r10_bench_n1.json 2> gpurun_out/r10_bench_n1.err
echo "bench rc=$?" >> gpurun_out/r10_status.txt
tail -3 gpurun_out/r10_gpu_tests.log; grep -i swiglu gpurun_out/r10_tune.log | head; head -c 400 gpurun_out/r10_bench_n1.json; cat gpurun_out/r10_status.txt
