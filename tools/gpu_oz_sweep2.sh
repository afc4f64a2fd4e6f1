#!/bin/bash
# re-sweep of the tile-order / look-ahead knobs with the dynamic tile scheduler: super-row height and SMs left to the panel chain
OUT=gpurun_out; mkdir -p $OUT
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-dmma --no-kernels > $OUT/bench_sw_$name.json 2> $OUT/bench_sw_$name.err
  python -c "
import json; d=json.load(open('$OUT/bench_sw_$name.json')); print('$name', round(d['value'],1), {k:round(v,2) for k,v in d['phases_ms'].items()}, round(d['roofline']['frac'],4), round(d['roofline']['isolated_launch']['ms'],3), d['clocks']['sm_mhz'])"
}
{
run base CK_OZ_SUPER_ROWS=16
run sr32 CK_OZ_SUPER_ROWS=32
run sr8 CK_OZ_SUPER_ROWS=8
run sr24 CK_OZ_SUPER_ROWS=24
run la6 CK_OZ_LA_SMS=6
run la16 CK_OZ_LA_SMS=16
run la0 CK_OZ_LA_SMS=0
run base2 CK_OZ_SUPER_ROWS=16
} 2>&1 | tee $OUT/oz_sweep2.log
