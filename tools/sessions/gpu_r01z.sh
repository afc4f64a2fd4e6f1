#!/bin/bash
TAG=${1:-r01z}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python tools/oz_probe.py --perf-only --sizes 38976x38976xL > $OUT/oz_perf_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ck_oz_gemm -s 2 -c 1 -o $OUT/prof_ozgemm_$TAG -f \
  python tools/oz_probe.py --perf-only --sizes 38976x38976xL > $OUT/ncu_ozgemm_$TAG.log 2>&1; echo "ncu_exit=$?"; tail -1 $OUT/oz_perf_$TAG.log | cut -c1-600
