SCFTB_MINB=4 python -m pytest tests/test_gpu_residual.py -m gpu -x -q 2>&1 | tail -2
for mb in 3 4; do SCFTB_MINB=$mb python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('MINB',$mb, d['value'], d['ms_per_step'], d['roofline']['frac'], d['problems_with_finite_residual_at_end'])"; done
