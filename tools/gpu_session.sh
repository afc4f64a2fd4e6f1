mkdir -p gpurun_out
L=gpurun_out/r02e_k4_variants.log; : > $L
V=$PWD/sif-xco2-cokriging_b200/cokrig_b200
for v in "" s3 s3c3 noskip; do
  echo "== variant '$v' recompute / gather" >> $L
  if [ -z "$v" ]; then unset COKRIG_B200_LIB; else export COKRIG_B200_LIB=$V/libvariant_$v.so; fi
  python tools/k4_run.py >> $L 2>&1; python tools/k4_run.py --gather >> $L 2>&1
done
cat $L
