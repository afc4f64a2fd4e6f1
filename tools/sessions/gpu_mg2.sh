#!/bin/bash
# 2-GPU session: multi-GPU parity + timing.  Usage: tools/gpu_mg2.sh <tag> [nproc]
TAG=${1:-r01d}
NP=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/smi_L_$TAG.txt
timeout 900 python -m pytest tests/test_gpu_parallel.py -m gpu -q > $OUT/pytest_gpu_parallel_$TAG.log 2>&1; echo "pytest_parallel_exit=$?"; tail -4 $OUT/pytest_gpu_parallel_$TAG.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29541 \
  tools/mg_check.py --points 20000 --targets 8833 --tile 1024 --out $OUT/mg_check_${NP}gpu_$TAG.json > $OUT/mg_check_${NP}gpu_$TAG.log 2>&1; echo "mg_check_exit=$?"; tail -3 $OUT/mg_check_${NP}gpu_$TAG.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29542 \
  bench.py --gpus $NP --steps 2 --warmup 3 > $OUT/bench_${NP}gpu_$TAG.json 2> $OUT/bench_${NP}gpu_$TAG.err; echo "bench_exit=$?"; cat $OUT/bench_${NP}gpu_$TAG.json | cut -c1-600
