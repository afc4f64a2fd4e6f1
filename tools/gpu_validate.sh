#!/bin/bash
# short single-GPU validation: GPU tests, smoke, both bench arms, optionally the ncu launch list of ONE bench step
# (the warm-up step's launches are skipped, not measured)
TAG=${1:-val}
OUT=gpurun_out
mkdir -p $OUT
SECONDS=0
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest_exit=$? t=$SECONDS"; tail -3 $OUT/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke_exit=$? t=$SECONDS"; tail -1 $OUT/smoke_$TAG.log
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench_exit=$? t=$SECONDS"; python -c "
import json; d=json.load(open('$OUT/bench_$TAG.json')); print(round(d['value'],1), d['phases_ms'], round(d['roofline']['frac'],3), d['e2e']['value'], d['gpu_launches'], d['clocks'], d['cpu_baseline']['value'])"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "bench_ref_exit=$? t=$SECONDS"; cut -c1-300 $OUT/bench_ref_$TAG.json
if [ "$2" = "ncu" ]; then
L=$(timeout 300 python bench.py --steps 1 --profile 2>&1 | sed -n 's/.*profile run: \([0-9]*\) launches.*/\1/p'); echo "launches per step: $L t=$SECONDS"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip ${L:-3486} --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 1 --profile > $OUT/ncu_launches_$TAG.log 2>&1; echo "ncu_launches_exit=$? t=$SECONDS"
fi
