#!/bin/bash
# 8-GPU validation: symmetric-memory collectives / fused kernels at world 8 (numerics + micro-benchmarks), Llama-3-8B bench,
# Mixtral-8x7B EP=8 bench
mkdir -p gpurun_out; rm -f gpurun_out/r11_status.txt
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/r11_gpus.txt; free -g >> gpurun_out/r11_gpus.txt
timeout 600 python -m pytest tests/gpu/test_symm_multi_gpu.py -x -q -s -k "8 or allgather_gemm" > gpurun_out/r11_symm_w8.log 2>&1
echo "symm w8 rc=$?" >> gpurun_out/r11_status.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
run() { name=$1; shift
  timeout 900 $TR --master-port $((29500 + RANDOM % 400)) bench.py --gpus 8 "$@" > gpurun_out/r11_$name.json 2> gpurun_out/r11_$name.err
  echo "$name rc=$?" >> gpurun_out/r11_status.txt; tail -c 1200 gpurun_out/r11_$name.json; echo; tail -3 gpurun_out/r11_$name.err | cut -c1-300; }
run l8b_native --steps 6 --warmup 3
run mixtral_ep8 --model mixtral-8x7b --zero-stage 2 --micro-batch 1 --steps 3 --warmup 3 --no-exposed
grep -E "passed|failed|error" gpurun_out/r11_symm_w8.log | tail -3; grep -E "roofline|all-gather|reduce-scatter|AG\+GEMM" gpurun_out/r11_symm_w8.log | cut -c1-400
cat gpurun_out/r11_status.txt
