mkdir -p gpurun_out
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_r1_${n}gpu.json 2> gpurun_out/bench_r1_${n}gpu.err
cut -c1-260 gpurun_out/bench_r1_${n}gpu.json; tail -2 gpurun_out/bench_r1_${n}gpu.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 tools/bench2d_multi.py 4095 4095 256 2>&1 | grep -E "world|Error|error" | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tests/mgpu_pcg2d.py 2>&1 | grep -E "MGPU|Error|error" | tail -3
