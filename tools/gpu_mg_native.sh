#!/bin/bash
# 2-GPU check of the C handle API (ck_mg_*): torchrun parity test, then the C3 system on both grid shapes with the native sweep timed beside the Python one
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 tools/mg_check.py --points 3000 --targets 1000 --tile 512 --grid 1x2 --native --out $OUT/mg_native_2gpu_small_1x2.json > $OUT/mg_native_small_1x2.log 2>&1; echo "small 1x2 exit=$?"; tail -2 $OUT/mg_native_small_1x2.log | cut -c1-1500
timeout 400 $TR --master-port 29542 tools/mg_check.py --points 20000 --targets 8833 --tile 1024 --grid 2x1 --native --steps 2 --skip-single --out $OUT/mg_native_2gpu_c3_2x1.json > $OUT/mg_native_c3_2x1.log 2>&1; echo "c3 2x1 exit=$?"; tail -2 $OUT/mg_native_c3_2x1.log | cut -c1-2500
