#!/bin/bash
# The A/B experiments of round 2 behind one entry point (each is what produced the like-named log under profiles/):
#
#   gpurun -- 'bash tools/gpu_experiments.sh <name>'            gpurun --gpus N -- '...' for the multi-GPU ones
#
#   l2hints     CK_OZ_L2_HINTS 0/3/1/2 on the C3 step + tools/kernel_bench.py          -> r02x_l2hints_sweep.log, r02x_kernel_bench_final.json
#   ncu_oz      ncu --set full of the largest C3 update of the INT8 kernel             -> r02*_ozgemm_*_ncu_full_summary.txt
#   dynamic     CK_OZ_DYNAMIC 1/0: tools/oz_probe.py (checks + isolated launches), C3 step -> r02y_oz_probe_*.json, r02y_oz_dynamic_sweep.log
#   knobs       CK_OZ_SUPER_ROWS / CK_OZ_LA_SMS on the C3 step                         -> r02y_oz_dynamic_knob_sweep.log
#   presplit    CK_OZ_PRESPLIT 1/0 on the C3 step                                      -> r02y_presplit_sweep.log
#   bench2      (2 GPUs) bench.py --gpus 2, dynamic vs static scheduler                -> bench_r02z_2gpu_*.json
#   panel_min   (2 GPUs) CK_MG_PANEL_MIN 40/24/16/32 on bench.py --gpus 2              -> r02z_mg_panel_min_sweep_2gpu.log
#   native N    (N = 2 or 4 GPUs) the C handle API vs the torch.distributed sweep      -> r02w_mg_native_*, r02x_mg_native_*
OUT=gpurun_out; mkdir -p $OUT
WHAT=${1:-help}
step() {  # name, env assignments...: one short C3 bench run, one summary line
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-dmma --no-kernels > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python -c "
import json; d=json.load(open('$OUT/bench_$name.json')); print('$name', round(d['value'],1), {k:round(v,2) for k,v in d['phases_ms'].items()}, round(d['roofline']['frac'],4), round(d['roofline']['isolated_launch']['ms'],3), d['clocks']['sm_mhz'])"
}
ncu_oz() {  # tag, env assignments...
  local tag=$1; shift
  env "$@" timeout 400 ncu --set full --clock-control none --import-source on -k regex:ck_oz_gemm_kernel --launch-skip 2 -c 1 \
    -o $OUT/prof_ozgemm_$tag -f python tools/oz_probe.py --perf-only --sizes 38976x38976xL > $OUT/ncu_oz_$tag.log 2>&1; echo "ncu $tag exit=$?"
  python tools/ncu_summary.py $OUT/prof_ozgemm_$tag.ncu-rep "ck_oz_gemm_kernel ($*): lower update rows=38976, K=1024 (the largest update of the C3 factorisation)" > $OUT/ozgemm_${tag}_ncu_full_summary.txt
  ncu -i $OUT/prof_ozgemm_$tag.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; d=rows[2]
for k in ('sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active','l1tex__m_xbar2l1tex_read_bytes.sum'):
    for i,n in enumerate(h):
        if n.startswith(k): print(n, rows[1][i], d[i])
" >> $OUT/ozgemm_${tag}_ncu_full_summary.txt
  grep -i "gpu__time_duration\|dram__bytes\|hit_rate\|imma\|xbar2l1tex_read_bytes.sum G" $OUT/ozgemm_${tag}_ncu_full_summary.txt
}
TR() { local secs=$1 np=$2 port=$3; shift 3; timeout $secs python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port "$@"; }
case $WHAT in
l2hints)
  for H in 0 3 1 2 0; do step l2hints_$H CK_OZ_L2_HINTS=$H; done 2>&1 | tee $OUT/l2hints_sweep.log
  timeout 600 python tools/kernel_bench.py --only k1,k2,k4,nll --out $OUT/kernels.json > $OUT/kernels.log 2>&1; echo "kernel_bench_exit=$?" ;;
ncu_oz)
  ncu_oz dyn1_h3 CK_OZ_DYNAMIC=1 CK_OZ_L2_HINTS=3; ncu_oz dyn0_h3 CK_OZ_DYNAMIC=0 CK_OZ_L2_HINTS=3; ncu_oz dyn0_h0 CK_OZ_DYNAMIC=0 CK_OZ_L2_HINTS=0 ;;
dynamic)
  timeout 300 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_parallel.py -q -x 2>&1 | tail -3
  for D in 1 0; do
    CK_OZ_DYNAMIC=$D timeout 300 python tools/oz_probe.py --perf --sizes 16384x16384xL,38976x38976xL,8832x32768xR --out $OUT/oz_probe_dyn$D.json > $OUT/oz_probe_dyn$D.log 2>&1; echo "probe dyn=$D exit=$?"
  done
  for D in 1 0 1; do step dyn$D CK_OZ_DYNAMIC=$D; done 2>&1 | tee $OUT/oz_dyn_sweep.log ;;
knobs)
  { step base CK_OZ_SUPER_ROWS=16; step sr32 CK_OZ_SUPER_ROWS=32; step sr8 CK_OZ_SUPER_ROWS=8; step sr24 CK_OZ_SUPER_ROWS=24
    step la6 CK_OZ_LA_SMS=6; step la16 CK_OZ_LA_SMS=16; step la0 CK_OZ_LA_SMS=0; step base2 CK_OZ_SUPER_ROWS=16; } 2>&1 | tee $OUT/oz_sweep2.log ;;
presplit)
  timeout 400 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_at_size.py tests/test_gpu_kernels.py -q -x 2>&1 | tail -3
  { step pre1 CK_OZ_PRESPLIT=1; step pre0 CK_OZ_PRESPLIT=0; step pre1b CK_OZ_PRESPLIT=1; step pre0b CK_OZ_PRESPLIT=0; } 2>&1 | tee $OUT/presplit_sweep.log ;;
bench2)
  TR 500 2 29561 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err; echo "bench2 exit=$?"
  CK_OZ_DYNAMIC=0 TR 300 2 29562 bench.py --gpus 2 --steps 5 --warmup 3 --no-extras > $OUT/bench_2gpu_static.json 2> $OUT/bench_2gpu_static.err; echo "static exit=$?" ;;
panel_min)
  for M in 40 24 16 32 40; do
    CK_MG_PANEL_MIN=$M TR 300 2 2957$((M%10)) bench.py --gpus 2 --steps 5 --warmup 3 --no-extras > $OUT/bench_2gpu_pmin$M.json 2> $OUT/bench_2gpu_pmin$M.err
    python -c "
import json; d=json.load(open('$OUT/bench_2gpu_pmin$M.json')); print('panel_min=$M', round(d['value'],1), round(d['ms_per_step'],2), round(d['native_handle_api']['ms_per_step'],2))"
  done 2>&1 | tee $OUT/mg_panel_min_sweep_2gpu.log ;;
native)
  NP=${2:-2}
  if [ "$NP" = 2 ]; then
    TR 300 2 29541 tools/mg_check.py --points 3000 --targets 1000 --tile 512 --grid 1x2 --native --out $OUT/mg_native_2gpu_small_1x2.json > $OUT/mg_native_small_1x2.log 2>&1; echo "small 1x2 exit=$?"
    TR 400 2 29542 tools/mg_check.py --points 20000 --targets 8833 --tile 1024 --grid 2x1 --native --steps 2 --skip-single --out $OUT/mg_native_2gpu_c3_2x1.json > $OUT/mg_native_c3_2x1.log 2>&1; echo "c3 2x1 exit=$?"
  else
    TR 400 4 29551 bench.py --gpus 4 --steps 3 --warmup 3 > $OUT/bench_4gpu.json 2> $OUT/bench_4gpu.err; echo "bench4 exit=$?"
    for G in 4x1 1x4; do
      TR 200 4 29552 tools/mg_check.py --points 4000 --targets 1500 --tile 512 --grid $G --native --skip-single --out $OUT/mg_native_4gpu_small_$G.json > $OUT/mg_native_4gpu_small_$G.log 2>&1; echo "small $G exit=$?"
    done
  fi ;;
*) sed -n 2,15p "$0" ;;
esac
