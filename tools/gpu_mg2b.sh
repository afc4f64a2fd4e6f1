#!/bin/bash
# 2-GPU session: block-cyclic sweep with INT8 updates vs FP64 DMMA.  Usage: tools/gpu_mg2b.sh <tag> [nproc]
TAG=${1:-r01o}
NP=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29541 \
  tools/mg_check.py --points 20000 --targets 8833 --tile 1024 --out $OUT/mg_check_${NP}gpu_$TAG.json > $OUT/mg_check_${NP}gpu_$TAG.log 2>&1; echo "mg_check_int8_exit=$?"; tail -2 $OUT/mg_check_${NP}gpu_$TAG.log | cut -c1-1500
CK_OZAKI=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29543 \
  tools/mg_check.py --points 20000 --targets 8833 --tile 1024 --skip-single --out $OUT/mg_check_${NP}gpu_${TAG}_dmma.json > $OUT/mg_check_${NP}gpu_${TAG}_dmma.log 2>&1; echo "mg_check_dmma_exit=$?"; tail -1 $OUT/mg_check_${NP}gpu_${TAG}_dmma.log | cut -c1-800
