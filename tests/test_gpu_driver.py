"""The C++ host driver (scft_b200/host/drivescft_b200.cpp) with the control flow of the reference's
mains: drivescft.cc:259-322 (read -> [broydn -> save -> refine] x levels) and 1D_FEM.c:289-370."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "scft_b200", "lib", "drivescft_b200")


@pytest.fixture(scope="module")
def sb():
    import scft_b200
    scft_b200.lib()
    if not os.path.exists(DRIVER):
        subprocess.run(["make", "-C", os.path.join(ROOT, "scft_b200", "csrc"), "all"], check=True)
    return scft_b200


def run_driver(args):
    p = subprocess.run([DRIVER] + args, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    rows = []
    for ln in p.stdout.splitlines():
        m = re.match(r"level (\d+): N=(\d+) check=(\d+) Error= (\S+) mean_field_free_energy=(\S+) Q=(\S+)", ln)
        if m:
            rows.append(dict(level=int(m[1]), N=int(m[2]), check=int(m[3]), err=float(m[4]), F=float(m[5]), Q=float(m[6])))
    return rows


def test_dealii_flow_two_refinement_levels(sb, oracle, fixtures, tmp_path):
    inp = str(tmp_path / "N=33_for_read.txt")
    sb.write_solution(inp, float(fixtures["n33_error"]), float(fixtures["n33_F"]), fixtures["n33_x"], fixtures["n33_eta"])
    rows = run_driver([inp, "--flow", "dealii", "--scheme", "irk4", "--solver", "broydn", "--levels", "3", "--tol", "1e-10",
                       "--outdir", str(tmp_path)])
    assert [r["N"] for r in rows] == [33, 65, 129]
    # level 0: the reference's own converged file; its recorded free energy
    assert rows[0]["err"] < 3e-9 and rows[0]["F"] == pytest.approx(float(fixtures["n33_F"]), abs=1e-11)
    for r in rows:
        assert r["err"] < 1e-8
    # free energy moves monotonically towards the fine-mesh value (Exp_m1024 header f = 1.9097e-3)
    assert rows[0]["F"] > rows[1]["F"] > rows[2]["F"] > float(fixtures["res1024_f"])
    # the saved files are in the reference format and re-evaluate to the reported residual on the ORACLE
    for r in rows:
        x, eta = sb.read_solution(str(tmp_path / f"solution_yita_1D_N={r['N']:03d}.txt"))
        assert len(x) == r["N"]
        ref = oracle.residual(oracle.eta_full(x, eta[1:-1]), oracle.f0_given(x), scheme=oracle.IRK4_CONSISTENT, nsteps=2048)
        assert np.abs(ref["out"]).max() < 1.5e-8
        assert oracle.free_energy(x, oracle.eta_full(x, eta[1:-1])) == pytest.approx(r["F"], abs=1e-11)


def test_dealii_flow_staged_anderson(sb, fixtures, tmp_path):
    inp = str(tmp_path / "N=33_for_read.txt")
    eta = fixtures["n33_eta"].copy()
    eta[1:-1] *= 1.02
    sb.write_solution(inp, 0.0, 0.0, fixtures["n33_x"], eta)
    rows = run_driver([inp, "--flow", "dealii", "--scheme", "ie", "--solver", "adm_chen", "--levels", "1", "--tol", "1e-8",
                       "--nsteps", "256", "--outdir", str(tmp_path)])
    assert rows[0]["N"] == 33 and rows[0]["err"] < 1e-7


def test_1dfem_flow(sb, fixtures, tmp_path):
    res = str(tmp_path / "Exp_m32_n2048_IE.res")
    with open(res, "w") as fh:
        fh.write("\n\nm = 32   n = 2048\n\nN=1000\nZ=1\n\n x/l phi eta phie phij\n-----\n")
        for a, b, c in zip(fixtures["res32_xl"], fixtures["res32_phi"], fixtures["res32_eta"]):
            fh.write(f" {a:.6e}  {b:.14e}  {c:.14e}  0.0  0.0\n")
    rows = run_driver([res, "--flow", "1dfem", "--outdir", str(tmp_path)])
    assert rows[0]["N"] == 33 and rows[0]["err"] < 1e-7     # broydn err = 1e-8 (1D_FEM.c:350)


def test_recheck_mode_reproduces_reference_file_error(sb, fixtures, tmp_path):
    """redoFxandshowError.cc:272-290: re-evaluate a saved solution and rewrite its files.  On the reference's own
    N=33_for_read.txt the driver must report the residual recorded in that file's header (1.42e-9) and its free energy,
    and write the detailed 2^18+1-point file (scft.cc:293-312)."""
    inp = str(tmp_path / "N=33_for_read.txt")
    sb.write_solution(inp, float(fixtures["n33_error"]), float(fixtures["n33_F"]), fixtures["n33_x"], fixtures["n33_eta"])
    rows = run_driver([inp, "--flow", "dealii", "--scheme", "irk4", "--recheck", "--detailed", "--outdir", str(tmp_path)])
    assert len(rows) == 1 and rows[0]["N"] == 33
    assert rows[0]["err"] < 2e-9 and rows[0]["F"] == pytest.approx(float(fixtures["n33_F"]), abs=1e-12)
    det = open(str(tmp_path / "detailedsolution_yita_1D_N=033.txt")).read().splitlines()
    assert len(det) == 2 + (1 << 18) + 1 and det[0].startswith("N= 33, ERROR=")
    i, xv, ev = det[2 + (1 << 17)].split(",")
    assert int(i) == 1 << 17 and float(xv) == pytest.approx(sb.L_REF / 2, abs=1e-12)
    assert float(ev) == pytest.approx(np.interp(sb.L_REF / 2, fixtures["n33_x"], fixtures["n33_eta"]), abs=2e-9)


def test_dealii_flow_preconditioned_mixing_to_m1024(sb, oracle, fixtures, tmp_path):
    """BASELINE.json configs[1]: the m=1024, n=2048 implicit-Euler hard-surface problem converged by Anderson mixing
    (preconditioned, pmixer.cu) through the reference's refinement flow; the final file re-evaluates on the oracle."""
    inp = str(tmp_path / "N=33_for_read.txt")
    sb.write_solution(inp, float(fixtures["n33_error"]), float(fixtures["n33_F"]), fixtures["n33_x"], fixtures["n33_eta"])
    rows = run_driver([inp, "--flow", "dealii", "--scheme", "ie_rowscale", "--solver", "padm", "--levels", "6", "--tol", "1e-9",
                       "--outdir", str(tmp_path)])
    assert [r["N"] for r in rows] == [33, 65, 129, 257, 513, 1025]
    assert all(r["check"] == 0 and r["err"] < 1e-9 for r in rows)
    x, eta = sb.read_solution(str(tmp_path / "solution_yita_1D_N=1025.txt"))
    ref = oracle.residual(oracle.eta_full(x, eta[1:-1]), oracle.f0_given(x), scheme=oracle.IE_ROWSCALE, nsteps=2048)
    assert np.abs(ref["out"]).max() < 2e-9       # the file keeps 15 decimals of eta
    assert rows[-1]["F"] == pytest.approx(0.001909977319, abs=2e-11)   # device Broyden's value, profiles/r1_continuation_m1024.txt


def test_dealii_flow_irk4_with_preconditioned_mixing(sb, oracle, fixtures, tmp_path):
    """the reference driver's own configuration (IRK4, drivescft.cc:130-146) with the preconditioned mixer instead of broydn: from
    the reference's converged N=33 file through three refinement levels; level 0 keeps the file's free energy, every saved file
    re-evaluates on the oracle"""
    inp = str(tmp_path / "N=33_for_read.txt")
    sb.write_solution(inp, float(fixtures["n33_error"]), float(fixtures["n33_F"]), fixtures["n33_x"], fixtures["n33_eta"])
    rows = run_driver([inp, "--flow", "dealii", "--scheme", "irk4", "--solver", "padm", "--levels", "3", "--tol", "1e-10",
                       "--outdir", str(tmp_path)])
    assert [r["N"] for r in rows] == [33, 65, 129] and all(r["check"] == 0 and r["err"] < 1e-10 for r in rows)
    assert rows[0]["F"] == pytest.approx(float(fixtures["n33_F"]), rel=2e-8)     # the file itself is converged to 1.4e-9 only
    for r in rows:
        x, eta = sb.read_solution(str(tmp_path / f"solution_yita_1D_N={r['N']:03d}.txt"))
        ref = oracle.residual(oracle.eta_full(x, eta[1:-1]), oracle.f0_given(x), scheme=oracle.IRK4_CONSISTENT, nsteps=2048)
        assert np.abs(ref["out"]).max() < 5e-10
        assert oracle.free_energy(x, oracle.eta_full(x, eta[1:-1])) == pytest.approx(r["F"], abs=1e-11)
