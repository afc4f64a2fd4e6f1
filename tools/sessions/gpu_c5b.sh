#!/bin/bash
# 8-GPU session: C5 with INT8 panel work; optional SM split.  Usage: tools/gpu_c5b.sh <tag> [nproc]
TAG=${1:-r01s}
NP=${2:-8}
OUT=gpurun_out
mkdir -p $OUT
for R in 0 40; do
CK_MG_PANEL_SMS=$R timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 2955$((R/40+1)) \
  tools/mg_check.py --workload c5 --targets 10000 --tile 1024 --steps 2 --skip-single --out $OUT/mg_c5_${NP}gpu_${TAG}_R$R.json > $OUT/mg_c5_${NP}gpu_${TAG}_R$R.log 2>&1; echo "c5_R${R}_exit=$?"; tail -1 $OUT/mg_c5_${NP}gpu_${TAG}_R$R.log | cut -c440-1100
done
