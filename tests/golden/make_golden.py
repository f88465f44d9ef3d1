"""Generate the committed golden fixtures under tests/golden/.

Run in the BUILD container only (needs /root/reference and oracle/_ref built from it):
    python tests/golden/make_golden.py

ref_fixtures.npz  — the reference's own data files, parsed (no source code):
    DEALII_SCFT/inputFiles/N=33_for_read.txt, Exp_m32_n2048_IE.res, Exp_m1024_n2048_IE.res,
    Matlab_files/inputFiles/solution_yita_1D_N= {43,59}.txt
ref_outputs.npz   — outputs of the UNMODIFIED reference C (oracle/_ref, compiled in place from
    /root/reference) on seeded inputs: romint, spline_chen(+gaussj), gaussj, adm_chen, adm, broydn.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def toyF(v):
    return np.array([np.cos(v[1]) - v[0], np.sin(v[0]) * 0.5 - v[1], 0.3 * v[0] - v[2] + 1])


def toy3(v):  # DEALII_SCFT/test_ADM.c:8-18
    x, y, z = v
    return np.array([x * y * z - 12., x * x + y * y - 8., x + y + z - 511.])


def toy2(v):  # 1D_FEM.c:372-379 (myfun)
    return np.array([v[0] * 0.5 - 2., v[1] * 0.5 + 3.])


def fp4(v):
    return np.array([np.cos(v[1]), 0.5 * np.sin(v[0]) + 0.1 * v[2], 0.3 * v[0] + 1, 0.2 * v[3] + 0.1 * v[0] * v[1]])


def main():
    O.build()
    fx = {}
    n33 = O.read_yita_file(f"{REF}/DEALII_SCFT/inputFiles/N=33_for_read.txt")
    fx.update(n33_x=n33["x"], n33_eta=n33["eta"], n33_error=n33["error"], n33_F=n33["F"])
    for m in (32, 1024):
        r = O.read_res_file(f"{REF}/Exp_m{m}_n2048_IE.res")
        fx[f"res{m}_xl"] = r["xl"]
        fx[f"res{m}_phi"] = r["phi"]
        fx[f"res{m}_eta"] = r["eta"]
        for k, v in r["meta"].items():
            fx[f"res{m}_{k}"] = v
    for n in (33, 43, 59):   # key matlab33s_*: "solution_yita_1D_N= 33.txt" (matlab33_* is solution_matlab_N=33)
        s = O.read_yita_file(f"{REF}/Matlab_files/inputFiles/solution_yita_1D_N= {n}.txt")
        if n == 33:
            fx["matlab33s_x"], fx["matlab33s_eta"] = s["x"], s["eta"]
            continue
        fx[f"matlab{n}_x"] = s["x"]
        fx[f"matlab{n}_eta"] = s["eta"]
    # a converged solution of Matlab_files/simple_FEM_1D_transient.m (drive_SCFT.m: tau=0.5302, L=3.72374, adm_chen to 1e-7):
    # the one reference artefact that pins the implicit-Euler / row-scaled scheme
    s = O.read_yita_file(f"{REF}/Matlab_files/inputFiles/solution_matlab_N=33")
    fx["matlab33_x"] = s["x"]
    fx["matlab33_eta"] = s["eta"]
    np.savez_compressed(os.path.join(OUT, "ref_fixtures.npz"), **fx)

    out = {}
    rng = np.random.default_rng(1234)
    f = rng.standard_normal(2049)
    out["romint_in"] = f
    out["romint_out"] = O.ref_romint(f, 1. / 2048)
    g = rng.standard_normal(65537)
    out["romint65537_in_seed"] = 1234
    out["romint65537_out"] = O.ref_romint(g, O.L_REF / 65536)
    # spline_chen: natural on the interior knots of a 33- and 129-node mesh incl. extrapolated ends,
    # not-a-knot transfer 33 -> 65 nodes as in refine_mesh (scft.cc:159-166)
    for N in (33, 129):
        x = O.mesh_uniform(N)
        y = rng.standard_normal(N - 2) * 5
        out[f"spline_nat{N}_y"] = y
        out[f"spline_nat{N}_yp"] = O.ref_spline(x[1:-1], y, x, O.SPLINE_NATURAL)
    x, xp = O.mesh_uniform(33), O.mesh_uniform(65)
    y = rng.standard_normal(31)
    out["spline_nak_y"] = y
    out["spline_nak_yp"] = O.ref_spline(x[1:-1], y, xp[1:-1], O.SPLINE_NOTAKNOT)
    xn = np.sort(np.concatenate([[0.0], rng.uniform(0.05, 3.6, 40), [O.L_REF]]))
    yn = rng.standard_normal(40)
    out["spline_nonuni_x"] = xn
    out["spline_nonuni_y"] = yn
    out["spline_nonuni_yp"] = O.ref_spline(xn[1:-1], yn, xn, O.SPLINE_NATURAL)
    A = rng.standard_normal((9, 9))
    B = rng.standard_normal((9, 2))
    rc, ai, bx = O.ref_gaussj(A, B)
    out.update(gaussj_A=A, gaussj_B=B, gaussj_Ainv=ai, gaussj_X=bx)
    rc, xs = O.ref_adm_chen(toyF, [1., 2., 3.], 1e-13, 500, 0.9, 3)
    out.update(admchen_toyF_rc=rc, admchen_toyF_x=xs)
    rc, xs = O.ref_adm_chen(toy3, [1., 2., 3.], 1e-12, 2000, 0.99, 30)
    out.update(admchen_toy3_rc=rc, admchen_toy3_x=xs)
    rc, xs = O.ref_adm(toy2, [1., 2.])
    out.update(adm_toy2_rc=rc, adm_toy2_x=xs)
    rc, xs = O.ref_adm(fp4, [1., 2., 3., 4.])
    out.update(adm_fp4_rc=rc, adm_fp4_x=xs)
    chk, xs, err, jc = O.ref_broydn(toyF, [1., 2., 3.], 1e-6)
    out.update(broydn_toyF_check=chk, broydn_toyF_x=xs, broydn_toyF_err=err, broydn_toyF_jc=jc)
    chk, xs, err, jc = O.ref_broydn(toy3, [1., 2., 3.], 1e-8)
    out.update(broydn_toy3_check=chk, broydn_toy3_x=xs, broydn_toy3_err=err, broydn_toy3_jc=jc)
    np.savez_compressed(os.path.join(OUT, "ref_outputs.npz"), **out)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
