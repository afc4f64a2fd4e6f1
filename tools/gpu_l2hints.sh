#!/bin/bash
# INT8 update kernel: L2 eviction-priority hints (CK_OZ_L2_HINTS bit 0: slices evict_last, bit 1: C evict_first) on the C3 step,
# then the per-kernel benches of the final build
OUT=gpurun_out; mkdir -p $OUT
for H in 0 3 1 2 0; do
  CK_OZ_L2_HINTS=$H timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-dmma --no-kernels > $OUT/bench_l2hints_$H.json 2> $OUT/bench_l2hints_$H.err
  python -c "
import json; d=json.load(open('$OUT/bench_l2hints_$H.json')); print('hints=$H', round(d['value'],1), {k:round(v,2) for k,v in d['phases_ms'].items()}, round(d['roofline']['frac'],4), d['roofline']['isolated_launch']['ms'], d['clocks']['sm_mhz'])"
done 2>&1 | tee $OUT/l2hints_sweep.log
timeout 600 python tools/kernel_bench.py --only k1,k2,k4,nll --out $OUT/kernels_r02x.json > $OUT/kernels_r02x.log 2>&1; echo "kernel_bench_exit=$?"
python -c "
import json; d=json.load(open('$OUT/kernels_r02x.json'))
for k,v in d.items():
    if isinstance(v, dict): print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if not isinstance(b,(list,dict))})"
