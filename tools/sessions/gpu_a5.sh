#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/kernel_bench.py --only k1 --quick --reps 1 > $OUT/k1_plain_r01aa.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ck_block_kernel -c 3 -o $OUT/prof_k1_r01aa -f \
  python tools/kernel_bench.py --only k1 --quick --reps 1 > $OUT/ncu_k1_r01aa.log 2>&1; echo "ncu_k1_exit=$?"
