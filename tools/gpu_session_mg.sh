# multi-GPU grid-shape sweep on NG GPUs
NG=${NG:-8}
mkdir -p gpurun_out
TAG=${TAG:-r02t}
L=gpurun_out/${TAG}_grid_${NG}gpu.log; : > $L
for gr in ${GRIDS:-4x2 2x4}; do
  echo "== grid $gr" >> $L
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 2 --warmup 3 --no-extras --grid $gr 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print(round(d['value'],1), round(d['ms_per_step'],1), d['phases_ms'])" >> $L 2>&1
done
cat $L
