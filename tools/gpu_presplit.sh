#!/bin/bash
# pre-split of the next panel on the side stream (CK_OZ_PRESPLIT) on the C3 step; parity tests of the factorisation paths first
OUT=gpurun_out; mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_at_size.py tests/test_gpu_kernels.py -q -x 2>&1 | tail -3
run() { local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-dmma --no-kernels > $OUT/bench_ps_$name.json 2> $OUT/bench_ps_$name.err
  python -c "
import json; d=json.load(open('$OUT/bench_ps_$name.json')); print('$name', round(d['value'],1), {k:round(v,2) for k,v in d['phases_ms'].items()}, round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"
}
{
run pre1 CK_OZ_PRESPLIT=1
run pre0 CK_OZ_PRESPLIT=0
run pre1_la16 CK_OZ_PRESPLIT=1 CK_OZ_LA_SMS=16
run pre1_la24 CK_OZ_PRESPLIT=1 CK_OZ_LA_SMS=24
run pre1b CK_OZ_PRESPLIT=1
run pre0b CK_OZ_PRESPLIT=0
} 2>&1 | tee $OUT/presplit_sweep.log
