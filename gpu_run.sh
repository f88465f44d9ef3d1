python - <<'PY'
import numpy as np, scft_b200 as sb
fx=np.load('tests/golden/ref_fixtures.npz')
sb.write_solution('/tmp/N33.txt', float(fx['n33_error']), float(fx['n33_F']), fx['n33_x'], fx['n33_eta'])
PY
./scft_b200/lib/drivescft_b200 /tmp/N33.txt --flow dealii --scheme ie_rowscale --solver adm_chen --levels 6 --tol 1e-9 --outdir /tmp | grep -E "^flow|^level"
./scft_b200/lib/drivescft_b200 /tmp/N33.txt --flow dealii --scheme ie --solver adm_chen --levels 6 --tol 1e-9 --outdir /tmp | grep -E "^flow|^level"
python -m pytest tests/test_gpu_edge_cases.py -m gpu -q 2>&1 | tail -3
