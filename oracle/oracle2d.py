"""CPU oracle of the 2-D path (numpy/scipy restatement; TEST INFRASTRUCTURE ONLY).

The reference's deal.II matrices on a structured Q1 mesh of nx x ny cells on [0,L] x [0,Ly]:
    A = (phi_i, phi_j), B = (grad phi_i, grad phi_j), C = (eta_h phi_i, phi_j)      scft.cc:643-656
with a 2x2 Gauss rule (QGauss<2>(2), scft.cc:610), homogeneous Dirichlet rows on x = 0 and x = L
(scft.cc:599-606), and an implicit-Euler step (A + ds (B + C)) q+ = A q — the SPD form a CG solver
can be applied to (the reference's own CG, scft.cc:698-705, is dead code and
its live stepper is the non-symmetric IRK4 block system).  Solved here with a sparse LU per field
update.  Density by Romberg weights over s (drivescft.cc:184-193).

DOF numbering: d = ix * (ny + 1) + iy.

Parity status: PINNED BY ORACLE ONLY — no reference artefact records a true 2-D run; the
y-invariant case is cross-checked against the 1-D oracle (which is fixture-pinned for IRK4).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl

from . import oracle as O

_G = np.array([-1.0, 1.0]) / np.sqrt(3.0)


def node_coords(nx, ny, L, Ly):
    x = L * np.arange(nx + 1) / nx
    y = Ly * np.arange(ny + 1) / ny
    return x, y


def assemble(nx, ny, L, Ly, eta):
    """dense-free COO assembly of A, B, C on all (nx+1)(ny+1) nodes (no constraints applied)"""
    x, y = node_coords(nx, ny, L, Ly)
    nd = (nx + 1) * (ny + 1)
    rows, cols, va, vb, vc = [], [], [], [], []
    for ex in range(nx):
        hx = x[ex + 1] - x[ex]
        for ey in range(ny):
            hy = y[ey + 1] - y[ey]
            dofs = [ex * (ny + 1) + ey, (ex + 1) * (ny + 1) + ey, ex * (ny + 1) + ey + 1, (ex + 1) * (ny + 1) + ey + 1]
            el_eta = eta[dofs]
            a = np.zeros((4, 4)); b = np.zeros((4, 4)); c = np.zeros((4, 4))
            for gx in _G:
                for gy in _G:
                    u, v = (gx + 1) / 2, (gy + 1) / 2
                    sh = np.array([(1 - u) * (1 - v), u * (1 - v), (1 - u) * v, u * v])
                    gr = np.array([[-(1 - v) / hx, -(1 - u) / hy], [(1 - v) / hx, -u / hy],
                                   [-v / hx, (1 - u) / hy], [v / hx, u / hy]])
                    jxw = hx * hy / 4
                    eq = float(sh @ el_eta)
                    a += np.outer(sh, sh) * jxw
                    b += (gr @ gr.T) * jxw
                    c += np.outer(sh, sh) * eq * jxw
            for i in range(4):
                for j in range(4):
                    rows.append(dofs[i]); cols.append(dofs[j])
                    va.append(a[i, j]); vb.append(b[i, j]); vc.append(c[i, j])
    mk = lambda v: sp.csr_matrix((v, (rows, cols)), shape=(nd, nd))
    return mk(va), mk(vb), mk(vc)


def free_dofs(nx, ny):
    return np.array([ix * (ny + 1) + iy for ix in range(1, nx) for iy in range(ny + 1)])


def residual(nx, ny, L, Ly, eta, tau=O.TAU_REF, nsteps=64, quadrature=O.QUAD_ROMBERG, sign=1.0, want_hist=False):
    """IE march on the 2-D mesh.  Returns dict(out, phi, Q[, hist]) on all nodes (walls: phi = 0)."""
    nd = (nx + 1) * (ny + 1)
    A, B, Cm = assemble(nx, ny, L, Ly, np.asarray(eta, dtype=np.float64))
    fr = free_dofs(nx, ny)
    dt = 1.0 / nsteps
    Af = A[fr][:, fr].tocsc()
    Tf = (A + dt * (B + Cm))[fr][:, fr].tocsc()
    lu = spl.splu(Tf)
    q = np.ones(len(fr))
    hist = np.zeros((len(fr), nsteps + 1))
    hist[:, 0] = q
    for j in range(1, nsteps + 1):
        q = lu.solve(Af @ q)
        hist[:, j] = q
    w = O.romberg_weights(nsteps, dt) if quadrature == O.QUAD_ROMBERG else np.r_[0.5, np.ones(nsteps - 1), 0.5] * dt
    phi = np.zeros(nd)
    phi[fr] = (hist * hist[:, ::-1]) @ w
    x, y = node_coords(nx, ny, L, Ly)
    f0x = O.f0_given(x, tau)
    f0 = np.repeat(f0x, ny + 1)
    out = sign * (f0 - phi)
    out[np.setdiff1d(np.arange(nd), fr)] = 0.0
    qfull = np.zeros(nd)
    qfull[fr] = q
    Q = float(np.ones(nd) @ (A @ qfull)) / (L * Ly)   # (1/|Omega|) int q(x,y,1) dx dy with the FEM mass matrix
    r = dict(out=out, phi=phi, Q=Q, f0=f0)
    if want_hist:
        h = np.zeros((nd, nsteps + 1))
        h[fr] = hist
        r["hist"] = h
    return r


def system_matrices(nx, ny, L, Ly, eta, nsteps):
    """(T, A) with the Dirichlet rows/columns replaced by identity / zero rows — the form the GPU stores"""
    nd = (nx + 1) * (ny + 1)
    A, B, Cm = assemble(nx, ny, L, Ly, np.asarray(eta, dtype=np.float64))
    dt = 1.0 / nsteps
    T = (A + dt * (B + Cm)).tolil()
    Al = A.tolil()
    wall = np.setdiff1d(np.arange(nd), free_dofs(nx, ny))
    for d in wall:
        T[d, :] = 0; T[:, d] = 0; T[d, d] = 1.0
        Al[d, :] = 0; Al[:, d] = 0
    return T.tocsr(), Al.tocsr()
