#!/bin/bash
# A/B session for the GEMM loader + look-ahead changes.  Usage: tools/gpu_ab.sh <tag>
TAG=${1:-r01c}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest_exit=$?"; tail -4 $OUT/pytest_gpu_$TAG.log
for cfg in "0 4" "1 4" "1 8" "1 2"; do
  set -- $cfg
  CK_LOOKAHEAD=$1 CK_AGG_BLOCKS=$2 timeout 600 python bench.py --steps 2 --no-cpu > $OUT/bench_${TAG}_la$1_agg$2.json 2> $OUT/bench_${TAG}_la$1_agg$2.err
  echo "la=$1 agg=$2 exit=$?"; python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_${TAG}_la$1_agg$2.json")); print(round(d["value"],1), {k:round(v,1) for k,v in d["phases_ms"].items()}, round(d["cholesky_TFs"],2), round(d["solve_TFs"],2), round(d["roofline"]["frac"],3))
except Exception as e: print("no json", e)
PY
done
timeout 600 python tools/kernel_bench.py --only k3,nll --out $OUT/kernels_k3_$TAG.json > $OUT/kernels_k3_$TAG.log 2>&1; echo "k3_exit=$?"
python -c "
import json; d=json.load(open('$OUT/kernels_k3_$TAG.json'))
for k,v in d.items():
    if k.startswith(('k3','nll')): print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items()})
print('cublas', d.get('cublas_dgemm_8192_TFs'))"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ck_gemm_nt -s 7 -c 1 -o $OUT/prof_gemm_$TAG -f \
  python tools/kernel_bench.py --only k3 --k3-sizes 16384 --reps 1 > $OUT/ncu_gemm_$TAG.log 2>&1; echo "ncu_gemm_exit=$?"
