#!/bin/bash
# Llama-3-70B ZeRO-3 + host offload (BASELINE config 3), full depth, 8 GPUs.  The 8-GPU box caps host memory at 1.1 TB (cgroup),
# full offload needs 1.13 TB of pinned state, so 70 % of the optimizer is host-stepped and 30 % device-stepped (Twin-Flow).
mkdir -p gpurun_out; rm -f gpurun_out/r16_status.txt
(echo "memory.max: $(cat /sys/fs/cgroup/memory.max 2>/dev/null)"; free -g; nproc) > gpurun_out/r16_host.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 1100 $TR --master-port 29611 bench.py --gpus 8 --model llama3-70b --offload cpu --zero-init --micro-batch 1 \
   --offload-ratio 0.7 --checkpoint-layers 40 --no-exposed --steps 2 --warmup 3 > gpurun_out/r16_l70b.json 2> gpurun_out/r16_l70b.err
echo "l70b rc=$?" >> gpurun_out/r16_status.txt
free -g >> gpurun_out/r16_host.txt
tail -c 1500 gpurun_out/r16_l70b.json; echo; tail -6 gpurun_out/r16_l70b.err | cut -c1-300; cat gpurun_out/r16_status.txt
