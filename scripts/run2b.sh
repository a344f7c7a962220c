#!/bin/bash
# GPU run 2b (2 GPUs): multi-GPU correctness + N=2 arms (native / hf / clip / gas) + reference arms for the ratios
mkdir -p gpurun_out; rm -f gpurun_out/r2b_status.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/gpu/test_symm_multi_gpu.py -x -q -s > gpurun_out/r2b_symm_tests.log 2>&1
echo "symm tests rc=$?" >> gpurun_out/r2b_status.txt
run() { # name, extra args
  name=$1; shift
  timeout 900 $TR --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 "$@" > gpurun_out/r2b_$name.json 2> gpurun_out/r2b_$name.err
  echo "$name rc=$?" >> gpurun_out/r2b_status.txt
}
run native --steps 8 --warmup 3
run native_clip --steps 6 --warmup 3 --clip 1.0
run native_gas4 --steps 3 --warmup 3 --gas 4
run hf --steps 6 --warmup 3 --model-impl hf
run ref --steps 6 --warmup 3 --impl reference
run ref_clip --steps 6 --warmup 3 --impl reference --clip 1.0
run ref_gas4 --steps 3 --warmup 3 --impl reference --gas 4
grep -E "passed|failed|RS\+Adam|all-gather|reduce-scatter|AG\+GEMM|MoE dispatch" gpurun_out/r2b_symm_tests.log | tail -12
for f in native native_clip native_gas4 hf ref ref_clip ref_gas4; do python - gpurun_out/r2b_$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d.get("value",0)), round(d.get("ms_per_step",0),1), round(d.get("e2e",{}).get("value",0)), d.get("exposed_comm"), d.get("unavailable"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done; cat gpurun_out/r2b_status.txt
