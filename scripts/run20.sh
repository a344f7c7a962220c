#!/bin/bash
# final verification on one GPU: full GPU test tier, smoke(), step profile, bench
mkdir -p gpurun_out; rm -f gpurun_out/r20_status.txt
timeout 900 python -m pytest tests/gpu -x -q -m gpu > gpurun_out/r20_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/r20_status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r20_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r20_status.txt
timeout 600 python scripts/profile_step.py --out gpurun_out/step_profile_r2f.txt > gpurun_out/r20_profile.log 2>&1
echo "profile rc=$?" >> gpurun_out/r20_status.txt
timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 > gpurun_out/r20_bench_n1.json 2> gpurun_out/r20_bench_n1.err
echo "bench rc=$?" >> gpurun_out/r20_status.txt
tail -4 gpurun_out/r20_gpu_tests.log; tail -2 gpurun_out/r20_smoke.log; head -45 gpurun_out/step_profile_r2f.txt | cut -c1-200; tail -c 600 gpurun_out/r20_bench_n1.json; cat gpurun_out/r20_status.txt
