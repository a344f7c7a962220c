#!/bin/bash
# 2-GPU validation of the Mixtral EP and 70B-offload bench arms at reduced depth (debug only: --layers invalidates the number)
mkdir -p gpurun_out; rm -f gpurun_out/r8_status.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { name=$1; shift
  timeout 900 $TR --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 "$@" > gpurun_out/r8_$name.json 2> gpurun_out/r8_$name.err
  echo "$name rc=$?" >> gpurun_out/r8_status.txt; tail -c 900 gpurun_out/r8_$name.json; echo; tail -5 gpurun_out/r8_$name.err | cut -c1-300; }
run mixtral_l2 --model mixtral-8x7b --layers 2 --zero-stage 2 --micro-batch 1 --steps 3 --warmup 3
run l70b_l4_offload --model llama3-70b --layers 4 --offload cpu --zero-init --micro-batch 1 --steps 2 --warmup 3
run l8b_twinflow --model llama3-8b --layers 8 --offload cpu --offload-ratio 0.5 --steps 2 --warmup 3
run l8b_native --steps 6 --warmup 3
cat gpurun_out/r8_status.txt
