#!/bin/bash
# One GPU-box session: tests, benches, ncu evidence.  Usage: tools/gpu_round.sh <tag>
TAG=${1:-r01b}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $OUT/smi_$TAG.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest_exit=$?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -5 $OUT/pytest_gpu_$TAG.log
timeout 900 python tools/kernel_bench.py --out $OUT/kernels_$TAG.json > $OUT/kernels_$TAG.log 2>&1; echo "kernel_bench_exit=$?"
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench_exit=$?"; cat $OUT/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "bench_ref_exit=$?"
# ncu: launch list of one bench step (after the bench above exited 0 without ncu)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 1 --profile > $OUT/ncu_launches_$TAG.log 2>&1; echo "ncu_launches_exit=$?"
# ncu --set full captures of each hot kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ck_gemm_nt -s 6 -c 2 -o $OUT/prof_gemm_$TAG -f \
  python tools/kernel_bench.py --only k3 --k3-sizes 16384 --reps 1 > $OUT/ncu_gemm_$TAG.log 2>&1; echo "ncu_gemm_exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ck_block_kernel -c 3 -o $OUT/prof_k1_$TAG -f \
  python tools/kernel_bench.py --only k1 --quick --reps 1 > $OUT/ncu_k1_$TAG.log 2>&1; echo "ncu_k1_exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ck_vario_bin_kernel|ck_vario_minmax_kernel" -c 2 -o $OUT/prof_k2_$TAG -f \
  python tools/kernel_bench.py --only k2 --quick --reps 1 > $OUT/ncu_k2_$TAG.log 2>&1; echo "ncu_k2_exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ck_local_predict_kernel|ck_potf2" -c 2 -o $OUT/prof_k4_$TAG -f \
  python tools/kernel_bench.py --only k4,k3 --k3-sizes 2048 --quick --reps 1 > $OUT/ncu_k4_$TAG.log 2>&1; echo "ncu_k4_exit=$?"
ls -la $OUT | tail -30
