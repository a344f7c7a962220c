#!/bin/bash
# 2 GPUs: ZeRO++ (qwZ / qgZ / hpZ) across GPUs, NVMe tier on the box disk, offload + Twin-Flow engine runs
mkdir -p gpurun_out; rm -f gpurun_out/r19_status.txt
timeout 900 python -m pytest tests/gpu/test_zeropp_multi_gpu.py -x -q -k "2 or nvme" > gpurun_out/r19_zeropp.log 2>&1
echo "zeropp rc=$?" >> gpurun_out/r19_status.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29631 bench.py --gpus 2 --model llama3-70b --layers 4 --offload cpu --zero-init --micro-batch 1 --no-exposed --steps 2 --warmup 3 > gpurun_out/r19_l70b_l4_n2.json 2> gpurun_out/r19_l70b_l4_n2.err
echo "l70b-l4 n2 rc=$?" >> gpurun_out/r19_status.txt
tail -5 gpurun_out/r19_zeropp.log; tail -c 500 gpurun_out/r19_l70b_l4_n2.json; cat gpurun_out/r19_status.txt
