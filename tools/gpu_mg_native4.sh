#!/bin/bash
# 4-GPU check of the C handle API: bench.py --gpus 4 (2x2 grid: the one-holder column exchange) and the 4x1 / 1x4 grids at a small size
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29551 bench.py --gpus 4 --steps 3 --warmup 3 > $OUT/bench_4gpu_r02x.json 2> $OUT/bench_4gpu_r02x.err; echo "bench4 exit=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_4gpu_r02x.json'))
print(d['value'], d['ms_per_step'], d['phases_ms'], d['e2e']['value'], d['native_handle_api'], d['parity_vs_single_gpu'], d['replicas'])
P
for G in 4x1 1x4; do
timeout 200 $TR --master-port 29552 tools/mg_check.py --points 4000 --targets 1500 --tile 512 --grid $G --native --skip-single --out $OUT/mg_native_4gpu_small_$G.json > $OUT/mg_native_4gpu_small_$G.log 2>&1; echo "small $G exit=$?"
python -c "
import json; d=json.load(open('$OUT/mg_native_4gpu_small_$G.json')); print({k:v for k,v in d.items() if 'native' in k or k in ('ok','solve_ms','grid')})"
done
