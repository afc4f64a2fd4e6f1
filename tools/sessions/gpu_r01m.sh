#!/bin/bash
TAG=${1:-r01m}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python tools/kernel_bench.py --only k1 --out $OUT/kernels_k1_$TAG.json > $OUT/kernels_k1_$TAG.log 2>&1; echo "k1_exit=$?"
python -c "
import json; d=json.load(open('$OUT/kernels_k1_$TAG.json'))
for k,v in d.items():
    if k.startswith('k1'): print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items()})"
timeout 300 python tools/kernel_bench.py --only k1 --quick --reps 1 > $OUT/k1_plain_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ck_block_kernel -c 3 -o $OUT/prof_k1_$TAG -f \
  python tools/kernel_bench.py --only k1 --quick --reps 1 > $OUT/ncu_k1_$TAG.log 2>&1; echo "ncu_k1_exit=$?"
