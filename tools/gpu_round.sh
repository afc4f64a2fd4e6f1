#!/bin/bash
# full single-GPU round: tests, smoke, benches (both arms), kernel benches, ncu launch list of one bench step
TAG=${1:-round}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest_exit=$?"; tail -3 $OUT/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke_exit=$?"; tail -1 $OUT/smoke_$TAG.log
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench_exit=$?"; python -c "
import json; d=json.load(open('$OUT/bench_$TAG.json')); print(round(d['value'],1), d['phases_ms'], round(d['roofline']['frac'],3), d['e2e']['value'], d['gpu_launches'], d['clocks'], d['cpu_baseline']['value'])"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "bench_ref_exit=$?"; cut -c1-300 $OUT/bench_ref_$TAG.json
timeout 900 python tools/kernel_bench.py --out $OUT/kernels_$TAG.json > $OUT/kernels_$TAG.log 2>&1; echo "kernel_bench_exit=$?"
python -c "
import json; d=json.load(open('$OUT/kernels_$TAG.json'))
for k,v in d.items():
    if isinstance(v, dict): print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if not isinstance(b,(list,dict))})"
for S in 1 8; do
  timeout 600 python tools/c4_bench.py --windows 64 --streams $S --out $OUT/c4_s${S}_$TAG.json > $OUT/c4_s${S}_$TAG.log 2>&1; echo "c4 streams=$S exit=$?"; tail -1 $OUT/c4_s${S}_$TAG.log | cut -c100-420
done
timeout 900 python bench.py --steps 1 --profile > $OUT/bench_profile_plain_$TAG.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 1 --profile > $OUT/ncu_launches_$TAG.log 2>&1; echo "ncu_launches_exit=$?"
