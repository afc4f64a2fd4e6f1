mkdir -p gpurun_out
T=r02o
R=/tmp/ncu_reps; mkdir -p $R
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ck_block_kernel -c 2 -o $R/k1 -f python tools/kernel_bench.py --only k1 --reps 1 > gpurun_out/${T}_ncu_k1.log 2>&1; echo "ncu_k1=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ck_vario -c 9 -o $R/k2 -f python tools/kernel_bench.py --only k2 --reps 1 > gpurun_out/${T}_ncu_k2.log 2>&1; echo "ncu_k2=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ck_local_predict -c 1 -o $R/k4 -f python tools/k4_run.py --gather --reps 1 --m 8000 > gpurun_out/${T}_ncu_k4.log 2>&1; echo "ncu_k4=$?"
for k in k1 k2 k4; do python tools/ncu_summary.py $R/$k.ncu-rep "round 2 ($T): $k" > gpurun_out/${T}_${k}_ncu_summary.txt; done
cat gpurun_out/${T}_k4_ncu_summary.txt | cut -c1-200
cat gpurun_out/${T}_k2_ncu_summary.txt | cut -c1-400 | head -30
du -sh gpurun_out
