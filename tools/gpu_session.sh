mkdir -p gpurun_out
L=gpurun_out/r02s_k4_shapes.log; : > $L
V=$PWD/sif-xco2-cokriging_b200/cokrig_b200
for v in "" c3 k32 k32c3 s3 s3c3 k32s3c2; do
  echo "== variant '$v' recompute / gather" >> $L
  if [ -z "$v" ]; then unset COKRIG_B200_LIB; else export COKRIG_B200_LIB=$V/libvariant_$v.so; fi
  timeout 120 python tools/k4_run.py >> $L 2>&1; timeout 120 python tools/k4_run.py --gather >> $L 2>&1
done
grep -E "==|targets_per_s" $L | sed -E 's/.*"targets_per_s": ([0-9.]+).*"gather": (true|false).*/\1 gather=\2/'
