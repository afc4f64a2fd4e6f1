#!/bin/bash
# r01h: INT8 tensor-core update path -- tests, K3 kernel bench, headline bench
TAG=${1:-r01h}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest_exit=$?"; tail -6 $OUT/pytest_gpu_$TAG.log
timeout 600 python tools/kernel_bench.py --only k3 --out $OUT/kernels_k3_$TAG.json > $OUT/kernels_k3_$TAG.log 2>&1; echo "k3_exit=$?"
python -c "
import json; d=json.load(open('$OUT/kernels_k3_$TAG.json'))
for k,v in d.items():
    if k.startswith('k3'): print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items()})"
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench_exit=$?"; tail -3 $OUT/bench_$TAG.err; python -c "
import json; d=json.load(open('$OUT/bench_$TAG.json')); print(round(d['value'],1), d['phases_ms'], round(d['roofline']['frac'],3), d['e2e']['value'], d['gpu_launches'])"
CK_OZAKI=0 timeout 900 python bench.py > $OUT/bench_${TAG}_dmma.json 2> $OUT/bench_${TAG}_dmma.err; echo "bench_dmma_exit=$?"; python -c "
import json; d=json.load(open('$OUT/bench_${TAG}_dmma.json')); print(round(d['value'],1), d['phases_ms'], round(d['roofline']['frac'],3), d['e2e']['value'], d['gpu_launches'])"
