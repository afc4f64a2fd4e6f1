#!/bin/bash
# 8-GPU session: C5 with the final kernels, 2x4 and 4x2 process grids.  Usage: tools/sessions/gpu_mg8c.sh <tag> [nproc]
TAG=${1:-r01af}
NP=${2:-8}
OUT=gpurun_out
mkdir -p $OUT
PORT=29580
for G in 2x4 4x2; do
PORT=$((PORT+1))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port $PORT \
  tools/mg_check.py --workload c5 --targets 10000 --tile 1024 --grid $G --steps 2 --skip-single --out $OUT/mg_c5_${NP}gpu_${TAG}_$G.json > $OUT/mg_c5_${NP}gpu_${TAG}_$G.log 2>&1; echo "c5_${G}_exit=$?"; tail -1 $OUT/mg_c5_${NP}gpu_${TAG}_$G.log | cut -c440-800
done
