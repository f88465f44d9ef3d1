timeout 500 python -m pytest tests/test_gpu_broyden_device.py -m gpu -q --tb=short 2>&1 | grep -E "^E  |passed|failed|Error" | cut -c1-220 | head -10
python - <<'PY'
import numpy as np, scft_b200 as sb
fx=np.load('tests/golden/ref_fixtures.npz')
sb.write_solution('/tmp/N33.txt', float(fx['n33_error']), float(fx['n33_F']), fx['n33_x'], fx['n33_eta'])
PY
./scft_b200/lib/drivescft_b200 /tmp/N33.txt --flow dealii --scheme ie_rowscale --solver broydn_dev --levels 6 --tol 1e-9 --outdir /tmp | grep -E "^flow|^level"
./scft_b200/lib/drivescft_b200 /tmp/N33.txt --flow dealii --scheme ie --solver broydn_dev --levels 6 --tol 1e-9 --outdir /tmp | grep -E "^flow|^level"
