#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -k "vario or dropin or parallel" > $OUT/pytest_a6.log 2>&1; echo "pytest_exit=$?"; tail -3 $OUT/pytest_a6.log
timeout 300 python tools/kernel_bench.py --only k2 --out $OUT/k2_a6.json > $OUT/k2_a6.log 2>&1; python -c "
import json; d=json.load(open('$OUT/k2_a6.json'))
for k,v in d.items():
    if k.startswith('k2'): print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items()})"
timeout 300 python tools/kernel_bench.py --only k2 --quick --reps 1 > $OUT/k2_plain_a6.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ck_vario_bin_kernel|ck_vario_minmax_kernel" -c 2 -o $OUT/prof_k2_r01ab -f \
  python tools/kernel_bench.py --only k2 --quick --reps 1 > $OUT/ncu_k2_a6.log 2>&1; echo "ncu_k2_exit=$?"
