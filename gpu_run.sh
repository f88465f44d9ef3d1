mkdir -p gpurun_out
timeout 280 python -m pytest tests/test_gpu_diblock.py -q -x 2>&1 | tail -25
