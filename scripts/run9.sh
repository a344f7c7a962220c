#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r9_status.txt
timeout 900 python -m pytest tests/gpu -x -q -m gpu > gpurun_out/r9_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/r9_status.txt
timeout 300 python scripts/bench_elementwise.py > gpurun_out/r9_elementwise.log 2>&1
echo "elementwise rc=$?" >> gpurun_out/r9_status.txt
timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 > gpurun_out/r9_bench_n1.json 2> gpurun_out/r9_bench_n1.err
echo "bench rc=$?" >> gpurun_out/r9_status.txt
for t in gemm attn wqtc symm; do
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_targets.py $t > gpurun_out/r9_memcheck_$t.log 2>&1
  echo "memcheck $t rc=$?" >> gpurun_out/r9_status.txt
done
for t in attn symm; do
  timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/sanitize_targets.py $t > gpurun_out/r9_racecheck_$t.log 2>&1
  echo "racecheck $t rc=$?" >> gpurun_out/r9_status.txt
done
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 7 python scripts/sanitize_targets.py symm > gpurun_out/r9_synccheck_symm.log 2>&1
echo "synccheck symm rc=$?" >> gpurun_out/r9_status.txt
tail -5 gpurun_out/r9_gpu_tests.log; tail -c 1500 gpurun_out/r9_bench_n1.json | head -c 1500; echo; for f in gpurun_out/r9_memcheck_*.log gpurun_out/r9_racecheck_*.log gpurun_out/r9_synccheck_*.log; do echo "== $f"; tail -4 $f; done; cat gpurun_out/r9_status.txt
