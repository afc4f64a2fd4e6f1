mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench_exit=$?"; tail -3 gpurun_out/r02f_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f_bench.json'))
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],1), d['phases_ms'])
r=d['roofline']; print('roofline frac', round(r['frac'],3), 'achieved', round(r['achieved'],1), 'peak', r['peak'], 'live int8', r.get('int8_gemm_live_TOPs'), 'frac live', r.get('frac_of_live_int8_sustained'))
print('isolated', r.get('isolated_launch'))
print('e2e', d['e2e']['value'], 'launches', d['gpu_launches'], d['clocks'])
print('dmma', d.get('value_fp64_dmma'))
for k,v in d.get('kernels',{}).items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a!='note'})
print('cpu', {k:v for k,v in d.get('cpu_baseline',{}).items() if k!='sample'})
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02f_bench_ref.json 2> gpurun_out/r02f_bench_ref.err; echo "bench_ref_exit=$?"; cut -c1-200 gpurun_out/r02f_bench_ref.json
