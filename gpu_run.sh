mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sweep_converge.py 32 1025 4 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -5 | tee gpurun_out/sweep_converge_2gpu.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 2>/dev/null | cut -c1-400
