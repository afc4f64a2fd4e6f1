#!/bin/bash
TAG=${1:-r01n}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest_exit=$?"; tail -5 $OUT/pytest_gpu_$TAG.log
timeout 600 python tools/kernel_bench.py --only k1 --out $OUT/kernels_k1_$TAG.json > $OUT/kernels_k1_$TAG.log 2>&1; echo "k1_exit=$?"
python -c "
import json; d=json.load(open('$OUT/kernels_k1_$TAG.json'))
for k,v in d.items():
    if k.startswith('k1'): print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items()})"
