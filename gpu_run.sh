mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3
python tools/lat.py 0 ab 2>&1 | tail -4 | tee gpurun_out/lat_ab.txt
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; cut -c1-300 gpurun_out/bench_r1_final.json
