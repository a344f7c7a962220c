#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r6_status.txt
timeout 120 python scripts/ncu_targets.py attn > gpurun_out/r6_plain.log 2>&1
echo "plain rc=$?" >> gpurun_out/r6_status.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -c 4 -o gpurun_out/attn_prof -f python scripts/ncu_targets.py attn > gpurun_out/r6_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/r6_status.txt
timeout 300 python -m pytest tests/gpu/test_inference_gpu.py -x -q -k "mixed_moe or wq_tc" > gpurun_out/r6_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r6_status.txt
tail -3 gpurun_out/r6_plain.log; tail -8 gpurun_out/r6_ncu.log; tail -5 gpurun_out/r6_tests.log; cat gpurun_out/r6_status.txt
