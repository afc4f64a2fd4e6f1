#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for SR in 4 8 16 32; do
  echo "== super rows $SR"
  CK_OZ_SUPER_ROWS=$SR timeout 300 python tools/oz_probe.py --perf-only --sizes 38976x38976xL,8832x32768xR 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['m'], d['n'], round(d['oz_gemm_ms'],3), round(d['oz_TFs_fp64_equiv'],1), {k:round(v,3) for k,v in d['cycles'].items() if 'frac' in k})"
done
CK_OZ_SUPER_ROWS=16 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:ck_oz_gemm -s 2 -c 1 python tools/oz_probe.py --perf-only --sizes 38976x38976xL 2>&1 | grep -E "dram__|gpu__time|lts__"
