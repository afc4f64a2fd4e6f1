#!/bin/bash
# round 2, session y: generic-nu table path of K1 (parity + throughput), L2 eviction hints of the INT8 update kernel
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_r02y.log 2>&1; echo "pytest_exit=$?"; tail -4 $OUT/pytest_gpu_r02y.log
timeout 300 python tools/kernel_bench.py --only k1,nll --out $OUT/r02y_k1.json > $OUT/r02y_k1.log 2>&1; echo "k1_exit=$?"; tail -12 $OUT/r02y_k1.log | cut -c1-400
for H in 0 1 2 3; do
  echo "== CK_OZ_L2_HINTS=$H"
  CK_OZ_L2_HINTS=$H timeout 200 python tools/oz_probe.py --perf-only --sizes 38976x38976xL,8832x38976xR 2>&1 | cut -c1-420
done
echo "== CK_OZ_L2_HINTS=3 CK_OZ_SUPER_ROWS=32"
CK_OZ_L2_HINTS=3 CK_OZ_SUPER_ROWS=32 timeout 200 python tools/oz_probe.py --perf-only --sizes 38976x38976xL 2>&1 | cut -c1-420
for H in 0 3; do
  CK_OZ_L2_HINTS=$H timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
    --clock-control none -k regex:ck_oz_gemm -c 2 --csv --log-file $OUT/r02y_oz_hints${H}_ncu.csv \
    python tools/oz_probe.py --perf-only --sizes 38976x38976xL > $OUT/r02y_oz_hints${H}_ncu.log 2>&1; echo "ncu_exit=$?"
  grep -v "^==" $OUT/r02y_oz_hints${H}_ncu.csv | cut -d, -f5,13- | tail -8
done
