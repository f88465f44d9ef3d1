timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tests/mgpu_pcg2d.py p2p 2>&1 | grep -E "MGPU|ScftError" | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 tools/bench2d_multi.py 4095 4095 64 p2p 2>&1 | grep -E "^world|ScftError" | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tools/bench2d_multi.py 1023 1023 2048 p2p 2>&1 | grep -E "^world|ScftError" | tail -3
