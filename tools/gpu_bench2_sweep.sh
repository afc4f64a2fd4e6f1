#!/bin/bash
# 2-GPU C3 sweep: smallest SM share of the panel stream (CK_MG_PANEL_MIN) now that the update kernel schedules its tiles dynamically
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for M in 40 24 16 32 40; do
CK_MG_PANEL_MIN=$M timeout 300 $TR --master-port 2957$((M%10)) bench.py --gpus 2 --steps 5 --warmup 3 --no-extras > $OUT/bench_2gpu_pmin$M.json 2> $OUT/bench_2gpu_pmin$M.err
python -c "
import json; d=json.load(open('$OUT/bench_2gpu_pmin$M.json')); print('panel_min=$M', round(d['value'],1), round(d['ms_per_step'],2), round(d['native_handle_api']['ms_per_step'],2))"
done 2>&1 | tee $OUT/mg_panel_min_sweep_2gpu.log
