#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r17_status.txt
timeout 600 python -m pytest tests/gpu/test_inference_gpu.py -x -q -k "wq or mixed" > gpurun_out/r17_wq_tests.log 2>&1
echo "wq tests rc=$?" >> gpurun_out/r17_status.txt
timeout 600 python -m pytest tests/gpu/test_attn_bias_gpu.py tests/gpu/test_fpdt_gpu.py -x -q > gpurun_out/r17_attn_bias_tests.log 2>&1
echo "attn_bias tests rc=$?" >> gpurun_out/r17_status.txt
timeout 600 python scripts/bench_attn_bias.py > gpurun_out/r17_attn_bias_bench.log 2>&1
echo "attn_bias bench rc=$?" >> gpurun_out/r17_status.txt
timeout 600 python scripts/bench_wq_tc.py > gpurun_out/r17_wq_tc_bench.log 2>&1
echo "wq bench rc=$?" >> gpurun_out/r17_status.txt
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k 'regex:^(fwd_kernel|bwd_dkdv_kernel|bwd_dq_kernel)$' -c 6 -o gpurun_out/ncu_attn_bias python scripts/ncu_targets.py evo > gpurun_out/r17_ncu_evo.log 2>&1
echo "ncu evo rc=$?" >> gpurun_out/r17_status.txt
timeout 600 $NCU -k regex:wq_tc -c 4 -o gpurun_out/ncu_wq_tc python scripts/ncu_targets.py wqtc > gpurun_out/r17_ncu_wqtc.log 2>&1
echo "ncu wqtc rc=$?" >> gpurun_out/r17_status.txt
for r in attn_bias wq_tc; do
  python scripts/ncu_summary.py gpurun_out/ncu_$r.ncu-rep > gpurun_out/ncu_${r}_summary.json 2> gpurun_out/ncu_${r}_summary.err
done
ncu -i gpurun_out/ncu_attn_bias.ncu-rep --page details --csv > gpurun_out/ncu_attn_bias_details.csv 2>/dev/null
python - <<'PY'
import json
c = json.load(open("examples/autotune_ds_config.json"))
c["autotuning"]["zero_stages"] = [2, 3]
json.dump(c, open("gpurun_out/autotune_cfg.json", "w"))
PY
PYTHONPATH=$PWD timeout 900 python -m deepspeed_b200.launcher.runner --autotuning tune --num_gpus 1 examples/autotune_train.py \
   --deepspeed_config gpurun_out/autotune_cfg.json > gpurun_out/r17_autotune.log 2>&1
echo "autotune rc=$?" >> gpurun_out/r17_status.txt
tail -3 gpurun_out/r17_wq_tests.log; tail -3 gpurun_out/r17_attn_bias_tests.log; grep '^{' gpurun_out/r17_attn_bias_bench.log | cut -c1-420; grep '^{' gpurun_out/r17_wq_tc_bench.log | cut -c1-260; tail -12 gpurun_out/r17_autotune.log | cut -c1-250; cat gpurun_out/r17_status.txt
