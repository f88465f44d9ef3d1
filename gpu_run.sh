mkdir -p gpurun_out
python -m pytest tests/test_gpu_broyden_device.py tests/test_gpu_sweep_converge.py -q 2>&1 | tail -5
(python tools/sweep_converge.py 64 1025 4 2>&1 | tail -4; python tools/sweep_converge.py 64 1025 8 2>&1 | tail -4) | tee gpurun_out/sweep_converge_threads.txt
