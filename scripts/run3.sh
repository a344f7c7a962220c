#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r3_status.txt
timeout 300 python -m pytest tests/gpu/test_attention_gpu.py -x -q > gpurun_out/r3_attn_tests.log 2>&1
echo "attn tests rc=$?" >> gpurun_out/r3_status.txt
timeout 300 python scripts/bench_attention.py > gpurun_out/r3_attn_bench.log 2>&1
echo "attn bench rc=$?" >> gpurun_out/r3_status.txt
timeout 600 python -m pytest tests/gpu/test_gemm_gpu.py tests/gpu/test_engine_gpu.py -x -q > gpurun_out/r3_tests.log 2>&1
echo "gemm+engine tests rc=$?" >> gpurun_out/r3_status.txt
timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 --model phi3-mini --zero-stage 2 --micro-batch 4 > gpurun_out/r3_phi3_n1.json 2> gpurun_out/r3_phi3_n1.err
echo "phi3 rc=$?" >> gpurun_out/r3_status.txt
timeout 900 python bench.py --gpus 1 --steps 8 --warmup 3 --model phi3-mini --zero-stage 2 --micro-batch 4 --impl reference > gpurun_out/r3_phi3_ref_n1.json 2> gpurun_out/r3_phi3_ref_n1.err
echo "phi3 ref rc=$?" >> gpurun_out/r3_status.txt
tail -5 gpurun_out/r3_attn_tests.log; tail -25 gpurun_out/r3_attn_bench.log; tail -3 gpurun_out/r3_tests.log; tail -c 600 gpurun_out/r3_phi3_n1.json; echo; tail -c 600 gpurun_out/r3_phi3_ref_n1.json; cat gpurun_out/r3_status.txt
