# multi-GPU session: strong-scaling bench (both arms) on NG GPUs; with C5=1 also the 1-degree global system (mg_check c5)
NG=${NG:-2}
mkdir -p gpurun_out
TAG=${TAG:-r02q}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_${NG}gpu.json 2> gpurun_out/${TAG}_bench_${NG}gpu.err; echo "bench_exit=$?"; tail -2 gpurun_out/${TAG}_bench_${NG}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_${NG}gpu.json').read().strip().splitlines()[-1])
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],1), d['phases_ms'], d['scaling'], d['config']['workload'][-60:])
print('e2e', d['e2e']['value'], 'replicas', d.get('replicas'), 'parity', d.get('parity_vs_single_gpu'), 'frac', d['roofline']['frac'])
PY
if [ "${C5:-0}" = "1" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 tools/mg_check.py --workload c5 --targets 10000 --tile 1024 --steps 2 --skip-single --out gpurun_out/${TAG}_mg_c5_${NG}gpu.json > gpurun_out/${TAG}_mg_c5_${NG}gpu.log 2>&1; echo "c5_exit=$?"; tail -1 gpurun_out/${TAG}_mg_c5_${NG}gpu.log | cut -c1-900
fi
if [ "${REF:-0}" = "1" ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $NG --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_ref_${NG}gpu.json 2>/dev/null; echo "ref_exit=$?"; cut -c1-250 gpurun_out/${TAG}_bench_ref_${NG}gpu.json
fi
