mkdir -p gpurun_out
python -m pytest tests/test_gpu_broyden_device.py tests/test_gpu_sweep_converge.py -q 2>&1 | tail -15
python tools/sweep_converge.py 32 2>&1 | tail -15 | tee gpurun_out/sweep_converge_1gpu.txt
