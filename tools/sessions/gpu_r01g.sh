#!/bin/bash
TAG=${1:-r01g}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest_exit=$?"; tail -4 $OUT/pytest_gpu_$TAG.log
timeout 600 python tools/kernel_bench.py --only k3 --out $OUT/kernels_k3_$TAG.json > $OUT/kernels_k3_$TAG.log 2>&1; echo "k3_exit=$?"
python -c "
import json; d=json.load(open('$OUT/kernels_k3_$TAG.json'))
for k,v in d.items():
    if k.startswith('k3'): print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items()})"
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench_exit=$?"; python -c "
import json; d=json.load(open('$OUT/bench_$TAG.json')); print(round(d['value'],1), d['phases_ms'], round(d['roofline']['frac'],3), d['e2e']['value'], d['gpu_launches'])"
for S in 1 2 4 8; do
  timeout 600 python tools/c4_bench.py --windows 64 --streams $S --out $OUT/c4_s${S}_$TAG.json > $OUT/c4_s${S}_$TAG.log 2>&1; echo "c4 streams=$S exit=$?"; tail -1 $OUT/c4_s${S}_$TAG.log | cut -c1-600
done
timeout 600 python tools/c4_bench.py --windows 4 --streams 2 --check --out $OUT/c4_check_$TAG.json > $OUT/c4_check_$TAG.log 2>&1; echo "c4 check exit=$?"; tail -1 $OUT/c4_check_$TAG.log | cut -c1-900
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ck_gemm_nt -s 31 -c 1 -o $OUT/prof_gemm_$TAG -f \
  python tools/kernel_bench.py --only k3 --k3-sizes 16384 --reps 1 > $OUT/ncu_gemm_$TAG.log 2>&1; echo "ncu_gemm_exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ck_potf2 -s 2 -c 1 -o $OUT/prof_potf2_$TAG -f \
  python tools/kernel_bench.py --only k3 --k3-sizes 2048 --reps 1 > $OUT/ncu_potf2_$TAG.log 2>&1; echo "ncu_potf2_exit=$?"
