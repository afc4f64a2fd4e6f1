#!/bin/bash
# dynamic tile scheduler of the INT8 update kernel (CK_OZ_DYNAMIC): correctness, isolated launches, the C3 step
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_parallel.py -q -x 2>&1 | tail -3
for D in 1 0; do
  CK_OZ_DYNAMIC=$D timeout 300 python tools/oz_probe.py --perf --sizes 16384x16384xL,38976x38976xL,8832x32768xR --out $OUT/oz_probe_dyn$D.json > $OUT/oz_probe_dyn$D.log 2>&1; echo "probe dyn=$D exit=$?"
  python -c "
import json; d=json.load(open('$OUT/oz_probe_dyn$D.json'))
print('cases max err', max(c['max_err_rel'] for c in d['cases']), all(c['pad_untouched'] for c in d['cases']))
for p in d['perf']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in p.items() if k!='cycles'}, {k:round(v,3) for k,v in p.get('cycles',{}).items() if 'frac' in k})"
done
for D in 1 0 1; do
  CK_OZ_DYNAMIC=$D timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-dmma --no-kernels > $OUT/bench_dyn$D.json 2> $OUT/bench_dyn$D.err
  python -c "
import json; d=json.load(open('$OUT/bench_dyn$D.json')); print('dyn=$D', round(d['value'],1), {k:round(v,2) for k,v in d['phases_ms'].items()}, round(d['roofline']['frac'],4), round(d['roofline']['isolated_launch']['ms'],3), d['clocks']['sm_mhz'])"
done 2>&1 | tee $OUT/oz_dyn_sweep.log
CK_OZ_DYNAMIC=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:ck_oz_gemm_kernel --launch-skip 2 -c 1 \
    -o $OUT/prof_ozgemm_r02_dyn -f python tools/oz_probe.py --perf-only --sizes 38976x38976xL > $OUT/ncu_oz_dyn.log 2>&1; echo "ncu exit=$?"
python tools/ncu_summary.py $OUT/prof_ozgemm_r02_dyn.ncu-rep "r02y ck_oz_gemm_kernel with the dynamic tile scheduler + L2 hints: lower update rows=38976, K=1024" > $OUT/r02y_ozgemm_dynamic_ncu_full_summary.txt
ncu -i $OUT/prof_ozgemm_r02_dyn.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; d=rows[2]
for k in ('sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active','l1tex__m_xbar2l1tex_read_bytes.sum'):
    for i,n in enumerate(h):
        if n.startswith(k): print(n, rows[1][i], d[i])
" >> $OUT/r02y_ozgemm_dynamic_ncu_full_summary.txt
grep -i "gpu__time_duration\|dram__bytes\|hit_rate\|imma\|xbar2l1tex_read_bytes.sum G" $OUT/r02y_ozgemm_dynamic_ncu_full_summary.txt
