mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r1_v4.json 2> gpurun_out/bench_r1_v4.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r1_v4_reference.json 2>> gpurun_out/bench_r1_v4.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_v4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:march_ie -s 3 -c 1 -f -o gpurun_out/prof_march_r1_v4 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:anderson -s 3 -c 1 -f -o gpurun_out/prof_anderson_r1_v4 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > gpurun_out/gpu_info.txt
nproc >> gpurun_out/gpu_info.txt; lscpu | grep "Model name" >> gpurun_out/gpu_info.txt
ls -la gpurun_out | tail -8
