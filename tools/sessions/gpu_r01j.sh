#!/bin/bash
TAG=${1:-r01j}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_ozaki.py -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest_exit=$?"; tail -4 $OUT/pytest_gpu_$TAG.log
timeout 300 python tools/oz_probe.py --perf-only > $OUT/oz_perf_$TAG.log 2>&1; echo "perf_exit=$?"; tail -4 $OUT/oz_perf_$TAG.log
