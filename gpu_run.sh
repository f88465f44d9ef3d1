python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_pcg2d.py -m gpu -q 2>&1 | tail -2
for b in 4 8; do SCFTB_2D_BLOCKS_PER_SM=$b python tools/bench2d.py 1023 1023 16 2>&1 | tail -1 | cut -c1-220; done
SCFTB_2D_BLOCKS_PER_SM=8 python tools/bench2d.py 1023 1023 2048 2>&1 | tail -1 | cut -c1-220
