mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_xcor.py -m gpu -q 2>&1 | tail -15
