NG=${NG:-2}
mkdir -p gpurun_out
TAG=${TAG:-r02k}
L=gpurun_out/${TAG}_sweep_${NG}gpu.log; : > $L
for mc in ${MCS:-8 32}; do
  echo "== CUDA_DEVICE_MAX_CONNECTIONS=$mc" >> $L
  CUDA_DEVICE_MAX_CONNECTIONS=$mc CK_MG_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 2 --warmup 3 --no-extras ${EXTRA:-} 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print(round(d['value'],1), round(d['ms_per_step'],1), d['phases_ms'])" >> $L 2>&1
  mv gpurun_out/mg_trace_rank0.json gpurun_out/${TAG}_mg_trace_${NG}gpu_mc${mc}_rank0.json
done
cat $L
