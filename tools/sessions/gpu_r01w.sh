#!/bin/bash
TAG=${1:-r01w}
OUT=gpurun_out
mkdir -p $OUT
for MR in 1024 2048 4096; do
  echo "== CK_OZ_MIN_ROWS=$MR"
  CK_OZ_MIN_ROWS=$MR timeout 300 python tools/kernel_bench.py --only k3 --k3-sizes 4096,6144,8192,12288,16384 --out $OUT/k3_${TAG}_mr$MR.json > $OUT/k3_${TAG}_mr$MR.log 2>&1
  python -c "
import json; d=json.load(open('$OUT/k3_${TAG}_mr$MR.json'))
for k,v in d.items():
    if k.startswith('k3'): print(k, round(v['potrf_ms'],2), round(v['potrf_TFs'],1), 'trsm', round(v['trsm_ms'],2), round(v['trsm_TFs'],1))"
  CK_OZ_MIN_ROWS=$MR timeout 300 python tools/c4_bench.py --windows 64 --streams 8 --out $OUT/c4_${TAG}_mr$MR.json > $OUT/c4_${TAG}_mr$MR.log 2>&1; tail -1 $OUT/c4_${TAG}_mr$MR.log | cut -c170-330
done
