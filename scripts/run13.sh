#!/bin/bash
# Llama-3-70B ZeRO-3 + host offload (BASELINE config 3) at 4 GPUs, full depth
mkdir -p gpurun_out; rm -f gpurun_out/r13_status.txt
free -g > gpurun_out/r13_host.txt; nproc >> gpurun_out/r13_host.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 1500 $TR --master-port 29611 bench.py --gpus 4 --model llama3-70b --offload cpu --zero-init --micro-batch 1 \
   --checkpoint-layers 40 --no-exposed --steps 2 --warmup 3 > gpurun_out/r13_l70b.json 2> gpurun_out/r13_l70b.err
echo "l70b rc=$?" >> gpurun_out/r13_status.txt
free -g >> gpurun_out/r13_host.txt
tail -c 1500 gpurun_out/r13_l70b.json; echo; tail -8 gpurun_out/r13_l70b.err | cut -c1-300; cat gpurun_out/r13_status.txt
timeout 900 python -m pytest tests/gpu/test_zeropp_multi_gpu.py -x -q > gpurun_out/r13_zeropp.log 2>&1
echo "zeropp rc=$?" >> gpurun_out/r13_status.txt
tail -5 gpurun_out/r13_zeropp.log; cat gpurun_out/r13_status.txt
