mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dropin.py tests/test_gpu_at_size.py -m gpu -q -k "vario" 2>&1 | tail -4
python tools/kernel_bench.py --only k2 --out gpurun_out/r02m_k2.json > gpurun_out/r02m_k2.log 2>&1; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02m_k2.json'))
for k,v in d.items():
    if k.startswith('k2'): print(k, v)
PY
