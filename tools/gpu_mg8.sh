#!/bin/bash
# 8-GPU session: C5 SM-split sweep, C4 windows sharded over the ranks, weak-scaling bench.  Usage: tools/gpu_mg8.sh <tag> [nproc]
TAG=${1:-r01y}
NP=${2:-8}
OUT=gpurun_out
mkdir -p $OUT
PORT=29560
for R in 24 32 48; do
PORT=$((PORT+1))
CK_MG_PANEL_SMS=$R timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port $PORT \
  tools/mg_check.py --workload c5 --targets 10000 --tile 1024 --steps 2 --skip-single --out $OUT/mg_c5_${NP}gpu_${TAG}_R$R.json > $OUT/mg_c5_${NP}gpu_${TAG}_R$R.log 2>&1; echo "c5_R${R}_exit=$?"; tail -1 $OUT/mg_c5_${NP}gpu_${TAG}_R$R.log | cut -c440-760
done
for S in 4 8; do
PORT=$((PORT+1))
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port $PORT \
  tools/c4_bench.py --windows 309 --streams $S --out $OUT/c4_${NP}gpu_${TAG}_s$S.json > $OUT/c4_${NP}gpu_${TAG}_s$S.log 2>&1; echo "c4_s${S}_exit=$?"; tail -1 $OUT/c4_${NP}gpu_${TAG}_s$S.log | cut -c100-420
done
PORT=$((PORT+1))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port $PORT \
  bench.py --gpus $NP --steps 2 --warmup 3 > $OUT/bench_${NP}gpu_$TAG.json 2> $OUT/bench_${NP}gpu_$TAG.err; echo "bench_exit=$?"; python -c "
import json; d=json.load(open('$OUT/bench_${NP}gpu_$TAG.json')); print(round(d['value'],1), d['phases_ms'], round(d['roofline']['frac'],3), d['e2e']['value'], d['clocks'])"
