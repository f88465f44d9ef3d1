mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:march_ie -c 1 -f -o gpurun_out/prof_march_r1_final python tools/prof_march.py one > gpurun_out/ncu_one.log 2>&1; tail -2 gpurun_out/ncu_one.log
ncu --set full --clock-control none --import-source on -k regex:march_ie -c 1 -f -o gpurun_out/prof_march_ab_r1 python tools/prof_march.py ab > gpurun_out/ncu_ab.log 2>&1; tail -2 gpurun_out/ncu_ab.log
ls -la gpurun_out/*.ncu-rep
