#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r5_status.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 4 -c 4 -o gpurun_out/attn_prof -f python scripts/ncu_targets.py attn > gpurun_out/r5_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/r5_status.txt
timeout 300 python -m pytest tests/gpu/test_inference_gpu.py -x -q -k "wq_tc" > gpurun_out/r5_wqtc_tests.log 2>&1
echo "wqtc tests rc=$?" >> gpurun_out/r5_status.txt
timeout 300 python -m pytest tests/gpu/test_inference_gpu.py tests/gpu/test_kernels_gpu.py -x -q > gpurun_out/r5_inf_tests.log 2>&1
echo "inf+kernel tests rc=$?" >> gpurun_out/r5_status.txt
tail -5 gpurun_out/r5_ncu.log; tail -15 gpurun_out/r5_wqtc_tests.log; tail -3 gpurun_out/r5_inf_tests.log; cat gpurun_out/r5_status.txt
