#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29561 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/bench_2gpu_r02z.json 2> $OUT/bench_2gpu_r02z.err; echo "bench2 exit=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_2gpu_r02z.json'))
print(d['value'], d['ms_per_step'], d['phases_ms'], d['e2e']['value'], d['native_handle_api'], d['parity_vs_single_gpu'], d['replicas'], d['roofline']['frac'])
P
CK_OZ_DYNAMIC=0 timeout 300 $TR --master-port 29562 bench.py --gpus 2 --steps 5 --warmup 3 --no-extras > $OUT/bench_2gpu_r02z_static.json 2> $OUT/bench_2gpu_r02z_static.err; echo "static exit=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu_r02z_static.json')); print('static', d['value'], d['ms_per_step'], d['native_handle_api'])"
