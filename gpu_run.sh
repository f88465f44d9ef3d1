mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -2
python tools/lat.py 0 2>&1 | tail -5
python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; cut -c1-200 gpurun_out/bench_r1_final.json
