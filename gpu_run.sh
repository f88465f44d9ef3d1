python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C8', d['value'], d['ms_per_step'], d['roofline']['frac'], d['problems_with_finite_residual_at_end'])"
