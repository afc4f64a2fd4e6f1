NG=${NG:-2}
mkdir -p gpurun_out
TAG=${TAG:-r02j}
CK_MG_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 2 --warmup 3 ${EXTRA:-} > gpurun_out/${TAG}_bench_${NG}gpu.json 2> gpurun_out/${TAG}_bench_${NG}gpu.err; echo "exit=$?"; tail -3 gpurun_out/${TAG}_bench_${NG}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_${NG}gpu.json').read().strip().splitlines()[-1])
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],1), d['phases_ms'], d['scaling'], d['config']['workload'][-60:])
print('e2e', d['e2e']['value'], 'replicas', d.get('replicas'), 'parity', d.get('parity_vs_single_gpu'))
PY
for r in 0 1; do mv gpurun_out/mg_trace_rank$r.json gpurun_out/${TAG}_mg_trace_${NG}gpu_rank$r.json; done
