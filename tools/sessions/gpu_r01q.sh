#!/bin/bash
TAG=${1:-r01q}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_parallel.py -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest_exit=$?"; tail -3 $OUT/pytest_gpu_$TAG.log
timeout 300 python tools/oz_probe.py --perf-only > $OUT/oz_perf_$TAG.log 2>&1; echo "perf_exit=$?"; cat $OUT/oz_perf_$TAG.log | cut -c1-420
timeout 900 python bench.py --no-cpu > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench_exit=$?"; python -c "
import json; d=json.load(open('$OUT/bench_$TAG.json')); print(round(d['value'],1), d['phases_ms'], round(d['roofline']['frac'],3), d['e2e']['value'], d['gpu_launches'], d['clocks'])"
