mkdir -p gpurun_out
L=gpurun_out/r02x_k4_tm128.log; : > $L
V=$PWD/sif-xco2-cokriging_b200/cokrig_b200
for v in "" tm128 tm128s3c2; do
  echo "== variant '$v' recompute / gather / gather k~720" >> $L
  if [ -z "$v" ]; then unset COKRIG_B200_LIB; else export COKRIG_B200_LIB=$V/libvariant_$v.so; fi
  timeout 120 python tools/k4_run.py >> $L 2>&1; timeout 120 python tools/k4_run.py --gather >> $L 2>&1; timeout 200 python tools/k4_run.py --gather --md 0.113 --reps 2 >> $L 2>&1
  timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_at_size.py -m gpu -x -q -k "local or point" 2>&1 | tail -1 >> $L
done
sed -E 's/.*"targets_per_s": ([0-9.]+).*"k_mean": ([0-9.]+).*"TFs": ([0-9.]+).*"gather": (true|false).*/k=\2 targets\/s=\1 TF=\3 gather=\4/' $L
