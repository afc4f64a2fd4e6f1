#!/bin/bash
TAG=${1:-r01x}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest_exit=$?"; tail -3 $OUT/pytest_gpu_$TAG.log
timeout 600 python tools/c4_bench.py --windows 4 --streams 2 --check --out $OUT/c4_check_$TAG.json > $OUT/c4_check_$TAG.log 2>&1; echo "c4 check exit=$?"; tail -1 $OUT/c4_check_$TAG.log | cut -c170-900
for MR in 512 1024; do
  echo "== CK_OZ_MIN_ROWS=$MR"
  CK_OZ_MIN_ROWS=$MR timeout 300 python tools/kernel_bench.py --only k3 --k3-sizes 2048,3072,4096,8192 --out $OUT/k3_${TAG}_mr$MR.json > $OUT/k3_${TAG}_mr$MR.log 2>&1
  python -c "
import json; d=json.load(open('$OUT/k3_${TAG}_mr$MR.json'))
for k,v in d.items():
    if k.startswith('k3'): print(k, round(v['potrf_ms'],2), round(v['potrf_TFs'],1), 'trsm', round(v['trsm_ms'],2), round(v['trsm_TFs'],1))"
  for S in 4 8 16; do CK_OZ_MIN_ROWS=$MR timeout 300 python tools/c4_bench.py --windows 64 --streams $S --out $OUT/c4_${TAG}_mr${MR}_s$S.json > $OUT/c4_${TAG}_mr${MR}_s$S.log 2>&1; tail -1 $OUT/c4_${TAG}_mr${MR}_s$S.log | cut -c150-300; done
done
