#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r24_status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r24_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r24_status.txt
timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 > gpurun_out/r24_bench_n1.json 2> gpurun_out/r24_bench_n1.err
echo "bench rc=$?" >> gpurun_out/r24_status.txt
timeout 900 python -m pytest tests/gpu -x -q -m gpu > gpurun_out/r24_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/r24_status.txt
tail -2 gpurun_out/r24_smoke.log; tail -c 500 gpurun_out/r24_bench_n1.json; tail -4 gpurun_out/r24_gpu_tests.log; tail -5 gpurun_out/r24_bench_n1.err | cut -c1-300; cat gpurun_out/r24_status.txt
