#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_ozaki.py -m gpu -q > $OUT/pytest_a7.log 2>&1; echo "pytest_exit=$?"; tail -3 $OUT/pytest_a7.log
for LA in 0 6 10 16; do
  echo "== CK_OZ_LA_SMS=$LA"
  CK_OZ_LA_SMS=$LA timeout 300 python tools/kernel_bench.py --only k3 --k3-sizes 16384,32768,40000 --out $OUT/k3_a7_la$LA.json > $OUT/k3_a7_la$LA.log 2>&1
  python -c "
import json; d=json.load(open('$OUT/k3_a7_la$LA.json'))
for k,v in d.items():
    if k.startswith('k3'): print(k, round(v['potrf_ms'],2), round(v['potrf_TFs'],1), 'trsm', round(v['trsm_ms'],2), round(v['trsm_TFs'],1), v['info'])"
done
for LA in 0 10; do
CK_OZ_LA_SMS=$LA timeout 600 python bench.py --no-cpu > $OUT/bench_a7_la$LA.json 2> $OUT/bench_a7_la$LA.err; echo "bench_exit=$?"; python -c "
import json; d=json.load(open('$OUT/bench_a7_la$LA.json')); print(round(d['value'],1), d['phases_ms'], round(d['roofline']['frac'],3), d['e2e']['value'], d['clocks'])"
done
