timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_pcg2d.py p2p 2>&1 | grep -E "MGPU|ScftError" | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/bench2d_multi.py 1023 1023 256 p2p 2>&1 | grep -E "^world|ScftError" | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tools/bench2d_multi.py 255 255 2048 p2p 2>&1 | grep -E "^world|ScftError" | tail -3
