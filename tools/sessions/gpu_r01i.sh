#!/bin/bash
TAG=${1:-r01i}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest_exit=$?"; tail -6 $OUT/pytest_gpu_$TAG.log
timeout 300 python tools/oz_probe.py --perf-only --sizes 32768x32768xL > $OUT/oz_perf_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ck_oz_gemm -s 2 -c 1 -o $OUT/prof_ozgemm_$TAG -f \
  python tools/oz_probe.py --perf-only --sizes 32768x32768xL > $OUT/ncu_ozgemm_$TAG.log 2>&1; echo "ncu_exit=$?"; tail -2 $OUT/oz_perf_$TAG.log
