mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dropin.py tests/test_gpu_at_size.py -m gpu -q -x 2>&1 | tail -3
python tools/kernel_bench.py --only k1 --out gpurun_out/r02u_k1.json > gpurun_out/r02u_k1.log 2>&1; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02u_k1.json'))
for k,v in d.items():
    if k.startswith('k1'): print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items()})
PY
python tools/k4_run.py --gather; python tools/k4_run.py
