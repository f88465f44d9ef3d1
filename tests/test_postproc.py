"""Host logic around the hot path (no GPU): spline with the reference's end conditions, mesh
refinement with not-a-knot transfer, result-file writer/reader — against the outputs of the
reference's own spline_chen (tests/golden/ref_outputs.npz) and the oracle."""
import os

import numpy as np
import pytest


@pytest.fixture(scope="module")
def sb():
    import scft_b200
    scft_b200.lib()
    return scft_b200


def test_spline_matches_reference_spline_chen(sb, oracle, refout):
    for N in (33, 129):
        x = oracle.mesh_uniform(N)
        yp = sb.spline(x[1:-1], refout[f"spline_nat{N}_y"], x, mode=0)
        assert np.abs(yp - refout[f"spline_nat{N}_yp"]).max() < 2e-13
    x, xp = oracle.mesh_uniform(33), oracle.mesh_uniform(65)
    yp = sb.spline(x[1:-1], refout["spline_nak_y"], xp[1:-1], mode=1)
    assert np.abs(yp - refout["spline_nak_yp"]).max() < 1e-12
    xn = refout["spline_nonuni_x"]
    yp = sb.spline(xn[1:-1], refout["spline_nonuni_y"], xn, mode=0)
    assert np.abs(yp - refout["spline_nonuni_yp"]).max() < 1e-11


def test_spline_given_second_derivative_and_errors(sb, oracle):
    rng = np.random.default_rng(0)
    x = np.sort(rng.uniform(0, 3, 12))
    y = rng.standard_normal(12)
    xp = np.linspace(-0.2, 3.2, 40)
    a = sb.spline(x, y, xp, mode=2, bc=0.7)
    b = oracle.spline(x, y, xp, oracle.SPLINE_GIVEN, 0.7)
    assert np.abs(a - b).max() < 1e-11
    with pytest.raises(sb.ScftError):
        sb.spline(x[:3], y[:3], xp, mode=1)     # not-a-knot needs > 3 points (spline_chen.c:44-49)


def test_refine_mesh_every_cell_cut_in_x(sb, oracle, fixtures):
    x = oracle.mesh_uniform(33)
    em = fixtures["n33_eta"][1:-1]
    xn, en = sb.refine_mesh(x, em)
    assert len(xn) == 65 and len(en) == 63
    assert np.array_equal(xn[::2], x) and np.allclose(xn[1::2], 0.5 * (x[:-1] + x[1:]), rtol=0, atol=1e-15)
    ref = oracle.spline(x[1:-1], em, xn[1:-1], oracle.SPLINE_NOTAKNOT)   # scft.cc:165
    assert np.abs(en - ref).max() < 1e-11
    assert np.abs(en[1::2] - em).max() < 1e-12                           # old nodes keep their values
    # five refinements reach the m=1024 mesh of BASELINE.json configs[1]: 33 -> 65 -> ... -> 1025
    N = 33
    for _ in range(5):
        x, em = sb.refine_mesh(x, em)
        N = 2 * N - 1
        assert len(x) == N
    assert N == 1025 and np.abs(x - oracle.mesh_uniform(1025)).max() < 1e-14


def test_solution_file_roundtrip_reference_format(sb, fixtures, tmp_path):
    p = str(tmp_path / "solution_yita_1D_N=033.txt")
    x, eta = fixtures["n33_x"], fixtures["n33_eta"]
    sb.write_solution(p, float(fixtures["n33_error"]), float(fixtures["n33_F"]), x, eta)
    lines = open(p).read().splitlines()
    assert lines[0] == "N= 33, ERROR= 1.422819e-09 "                      # scft.cc:328-329
    assert lines[1] == "mean_field_free_energy, 0.001945037280931 "       # scft.cc:330
    assert lines[2] == "0,0.000000000000000,1.011259322027922"            # first row of the reference file
    xr, er = sb.read_solution(p)
    assert np.abs(xr - x).max() < 1e-15 and np.abs(er - eta).max() < 1e-15
    # the oracle-side reader of the same format agrees
    from oracle import oracle as O
    d = O.read_yita_file(p)
    assert d["N"] == 33 and np.array_equal(d["eta"], er)


def test_res_reader(sb, fixtures, tmp_path):
    p = str(tmp_path / "Exp.res")
    with open(p, "w") as fh:
        fh.write("m = 32   n = 2048\n\nN=1000  l=3.7\nZ=1 f=2\n\n     x/l  phi eta phie phij\n---- ----\n")
        fh.write("\n\n")
        for a, b, c in zip(fixtures["res32_xl"], fixtures["res32_phi"], fixtures["res32_eta"]):
            fh.write(f" {a:.6e}  {b:.14e}  {c:.14e}  0.0  0.0\n")
    xl, phi, eta = sb.read_res(p, 33)
    assert np.allclose(eta, fixtures["res32_eta"], rtol=1e-14, atol=0) and np.allclose(phi, fixtures["res32_phi"], rtol=1e-14)


def test_adaptive_refinement_follows_matlab_prototype(sb, oracle, fixtures):
    """Matlab_files/refine_mesh.m: cells with |d eta/dx| >= 10 x median and the two wall cells are bisected.
    The prototype's own mesh sequence 33 -> 43 -> 59 (Matlab_files/inputFiles/solution_yita_1D_N= 43/59.txt) is the
    result of exactly this rule applied to its converged fields; the node COUNT after one refinement of the
    N=33 deal.II field must be in that range and the new nodes must be midpoints of flagged cells."""
    x = oracle.mesh_uniform(33)
    em = fixtures["n33_eta"][1:-1]
    xn, en = sb.refine_mesh_adaptive(x, em)
    assert 35 <= len(xn) <= 65 and len(en) == len(xn) - 2
    assert np.all(np.diff(xn) > 0) and xn[0] == x[0] and xn[-1] == x[-1]
    assert set(np.round(x, 12)).issubset(set(np.round(xn, 12)))            # old nodes are kept
    new = np.array(sorted(set(np.round(xn, 12)) - set(np.round(x, 12))))
    mids = 0.5 * (x[:-1] + x[1:])
    assert all(np.abs(mids - v).min() < 1e-12 for v in new)               # new nodes are cell midpoints
    assert abs(new[0] - mids[0]) < 1e-12 and abs(new[-1] - mids[-1]) < 1e-12   # wall cells always cut
    # flagged cells have the steepest field
    grad = np.abs(np.diff(em) / np.diff(x[1:-1]))
    thr = 10 * np.median(np.r_[np.inf, grad, np.inf])
    expect = [mids[0]] + [mids[c + 1] for c in range(len(grad)) if grad[c] >= thr] + [mids[-1]]
    assert np.allclose(new, expect, rtol=0, atol=1e-12)
    ref = oracle.spline(x[1:-1], em, xn[1:-1], oracle.SPLINE_NOTAKNOT)
    assert np.abs(en - ref).max() < 1e-11
    # the MATLAB-era adaptive meshes in the reference's fixtures have the same structure: refined near the walls
    for n_nodes in (43, 59):
        xm = fixtures[f"matlab{n_nodes}_x"]
        h = np.diff(xm)
        assert h[:3].max() < h[len(h) // 2]


def test_adaptive_refinement_reproduces_the_prototypes_own_mesh(sb, fixtures):
    """reference artefact: Matlab_files/refine_mesh.m applied to the field of 'solution_yita_1D_N= 33.txt' produced the
    43-node mesh stored in 'solution_yita_1D_N= 43.txt' — scftb_refine_mesh_adaptive gives the same nodes"""
    x, eta = fixtures["matlab33s_x"], fixtures["matlab33s_eta"]
    xn, en = sb.refine_mesh_adaptive(x, eta[1:-1])
    assert len(xn) == 43 and np.abs(xn - fixtures["matlab43_x"]).max() < 1e-12
    assert len(en) == 41 and np.all(np.isfinite(en))


def test_detailed_solution_file_layout(sb, tmp_path):
    """detailedsolution_yita_1D_N=<N>.txt (scft.cc:293-312): header as the solution file, then i,x_i,eta_h(x_i) on
    equidistant points with eta_h piecewise linear on the mesh (FEFieldFunction on Q1)"""
    x = np.array([0.0, 0.5, 1.5, 2.0])
    eta = np.array([1.0, 3.0, -1.0, 0.0])
    path = str(tmp_path / "detailedsolution_yita_1D_N=004.txt")
    sb.write_detailed_solution(path, 1.5e-9, 0.00194, x, eta, nplot=9)
    lines = open(path).read().splitlines()
    assert lines[0] == "N= 4, ERROR= 1.500000e-09 " and lines[1] == "mean_field_free_energy, 0.001940000000000 "
    assert len(lines) == 2 + 9
    xs = np.array([float(ln.split(",")[1]) for ln in lines[2:]])
    vs = np.array([float(ln.split(",")[2]) for ln in lines[2:]])
    assert np.allclose(xs, np.linspace(0, 2, 9), atol=1e-15) and np.allclose(vs, np.interp(xs, x, eta), atol=1e-14)
    assert lines[2].startswith("0,0.000000000000000,1.000000000000000")
