#!/bin/bash
# 1 GPU: attention-bias micro-benchmark, ncu captures of the round-2 kernels, a real autotuning run, full GPU test tier, bench
mkdir -p gpurun_out profiles; rm -f gpurun_out/r14_status.txt
(echo "memory.max: $(cat /sys/fs/cgroup/memory.max 2>/dev/null)"; echo "memory.high: $(cat /sys/fs/cgroup/memory.high 2>/dev/null)";
 echo "memlock: $(ulimit -l)"; echo "nproc: $(nproc)"; free -g; df -h /tmp /dev/shm | tail -3; nvidia-smi -L | wc -l) > gpurun_out/r14_box.txt 2>&1
timeout 600 python scripts/bench_attn_bias.py > gpurun_out/r14_attn_bias_bench.log 2>&1
echo "attn_bias bench rc=$?" >> gpurun_out/r14_status.txt
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:gemm_2cta -c 4 -o gpurun_out/ncu_mlp_epilogues python scripts/ncu_targets.py mlp > gpurun_out/r14_ncu_mlp.log 2>&1
echo "ncu mlp rc=$?" >> gpurun_out/r14_status.txt
timeout 600 $NCU -k 'regex:battn' -c 8 -o gpurun_out/ncu_attn_bias python scripts/ncu_targets.py evo > gpurun_out/r14_ncu_evo.log 2>&1
echo "ncu evo rc=$?" >> gpurun_out/r14_status.txt
timeout 600 $NCU -k regex:wq_tc -c 4 -o gpurun_out/ncu_wq_tc python scripts/ncu_targets.py wqtc > gpurun_out/r14_ncu_wqtc.log 2>&1
echo "ncu wqtc rc=$?" >> gpurun_out/r14_status.txt
timeout 600 $NCU -k 'regex:attn_(fwd|bwd)' -c 8 -o gpurun_out/ncu_attn_sm100 python scripts/ncu_targets.py attn > gpurun_out/r14_ncu_attn.log 2>&1
echo "ncu attn rc=$?" >> gpurun_out/r14_status.txt
for r in mlp_epilogues attn_bias wq_tc attn_sm100; do
  python scripts/ncu_summary.py gpurun_out/ncu_$r.ncu-rep > gpurun_out/ncu_${r}_summary.json 2> gpurun_out/ncu_${r}_summary.err
done
# autotuner end to end on the box: real launcher subprocesses, engine-side measurement, boosted-trees tuner
python - <<'PY'
import json
c = json.load(open("examples/autotune_ds_config.json"))
c["autotuning"]["zero_stages"] = [2, 3]
json.dump(c, open("gpurun_out/autotune_cfg.json", "w"))
PY
timeout 900 python -m deepspeed_b200.launcher.runner --autotuning tune --num_gpus 1 examples/autotune_train.py \
   --deepspeed_config gpurun_out/autotune_cfg.json > gpurun_out/r14_autotune.log 2>&1
echo "autotune rc=$?" >> gpurun_out/r14_status.txt
timeout 900 python -m pytest tests/gpu -x -q -m gpu > gpurun_out/r14_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/r14_status.txt
timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 > gpurun_out/r14_bench_n1.json 2> gpurun_out/r14_bench_n1.err
echo "bench rc=$?" >> gpurun_out/r14_status.txt
tail -12 gpurun_out/r14_attn_bias_bench.log | cut -c1-600; tail -15 gpurun_out/r14_autotune.log | cut -c1-250; tail -4 gpurun_out/r14_gpu_tests.log; cat gpurun_out/r14_status.txt
