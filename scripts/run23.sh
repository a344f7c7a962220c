#!/bin/bash
# N=2, both arms back to back on one box (final tree)
mkdir -p gpurun_out; rm -f gpurun_out/r23_status.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 800 $TR --master-port 29641 bench.py --impl reference --gpus 2 --steps 8 --warmup 3 > gpurun_out/r23_ref_n2.json 2> gpurun_out/r23_ref_n2.err
echo "ref rc=$?" >> gpurun_out/r23_status.txt
timeout 600 $TR --master-port 29642 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r23_ours_n2.json 2> gpurun_out/r23_ours_n2.err
echo "ours rc=$?" >> gpurun_out/r23_status.txt
tail -c 400 gpurun_out/r23_ref_n2.json; echo; tail -c 400 gpurun_out/r23_ours_n2.json; cat gpurun_out/r23_status.txt
