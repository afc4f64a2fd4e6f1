#!/bin/bash
# ncu --set full of ONE launch of the INT8 update kernel on the largest C3 update (lower, rows = 38976, K = 1024), with and without the L2 hints
OUT=gpurun_out; mkdir -p $OUT
for H in 3 0; do
  CK_OZ_L2_HINTS=$H timeout 200 python tools/oz_probe.py --perf-only --sizes 38976x38976xL > $OUT/oz_probe_plain_h$H.log 2>&1 || { echo "plain run failed h=$H"; tail -5 $OUT/oz_probe_plain_h$H.log; continue; }
  tail -1 $OUT/oz_probe_plain_h$H.log | cut -c1-300
  CK_OZ_L2_HINTS=$H timeout 400 ncu --set full --clock-control none --import-source on -k regex:ck_oz_gemm_kernel --launch-skip 2 -c 1 \
    -o $OUT/prof_ozgemm_r02_h$H -f python tools/oz_probe.py --perf-only --sizes 38976x38976xL > $OUT/ncu_oz_h$H.log 2>&1; echo "ncu h=$H exit=$?"
  python tools/ncu_summary.py $OUT/prof_ozgemm_r02_h$H.ncu-rep "r02 ck_oz_gemm_kernel, CK_OZ_L2_HINTS=$H: lower update rows=38976, K=1024 (the largest update of the C3 factorisation)" > $OUT/r02_ozgemm_h${H}_ncu_full_summary.txt
  ncu -i $OUT/prof_ozgemm_r02_h$H.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; d=rows[2]
for k in ('sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor_subpipe_imma.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','l1tex__m_xbar2l1tex_read_bytes.sum','lts__t_bytes.sum','lts__t_sectors_srcunit_tex_op_read.sum','sm__cycles_elapsed.avg.per_second','gpc__cycles_elapsed.avg.per_second'):
    for i,n in enumerate(h):
        if n.startswith(k): print(n, rows[1][i], d[i])
" >> $OUT/r02_ozgemm_h${H}_ncu_full_summary.txt
  grep -i "gpu__time_duration\|dram__bytes\|hit_rate\|tensor\|xbar" $OUT/r02_ozgemm_h${H}_ncu_full_summary.txt
done
