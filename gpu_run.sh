timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_pcg2d.py 2>&1 | grep -E "rank|MGPU|Error|error" | tail -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/bench2d_multi.py 1023 1023 256 2>&1 | grep -E "world|Error|error" | tail -3
python tools/bench2d.py 1023 1023 256 2>&1 | tail -1
