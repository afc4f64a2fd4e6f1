#!/bin/bash
# 8-GPU session: BASELINE config 5 (1 degree global bivariate cokriging, N = 129 600) + C3 strong scaling point.
TAG=${1:-r01f}
NP=${2:-8}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/smi_L_$TAG.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29551 \
  tools/mg_check.py --workload c5 --targets 10000 --tile 1024 --steps 2 --skip-single --out $OUT/mg_c5_${NP}gpu_$TAG.json > $OUT/mg_c5_${NP}gpu_$TAG.log 2>&1; echo "c5_exit=$?"; tail -2 $OUT/mg_c5_${NP}gpu_$TAG.log | cut -c1-1800
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29552 \
  tools/mg_check.py --points 20000 --targets 8833 --tile 1024 --steps 2 --skip-single --out $OUT/mg_c3_${NP}gpu_$TAG.json > $OUT/mg_c3_${NP}gpu_$TAG.log 2>&1; echo "c3_exit=$?"; tail -1 $OUT/mg_c3_${NP}gpu_$TAG.log | cut -c1-1200
