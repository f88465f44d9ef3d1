"""Parameter sweeps of independent SCFT problems and their sharding across ranks.

SURVEY.md §8(d) item 3 / §8(e): problem p of the sweep lives on a (tau, L) grid cell with a seeded
perturbation of the initial field; problems are independent, so rank r of G takes the contiguous
block [p0, p1) and there is NO collective on the data path — only one all-gather of per-problem
scalars (Q, error norm, iteration count) at the end.  One process per GPU; torch.distributed is
plumbing (NCCL on the GPU box, gloo in the CPU tests).
"""
import os

import numpy as np

TAU_GRID = np.linspace(0.40, 0.66, 16)
L_GRID = np.linspace(3.2, 4.2, 16)


def sweep_params(p):
    """(tau, L, seed) of sweep problem p: 16 x 16 (tau, L) cells, then the seed axis."""
    cell, _ = p % 256, p // 256
    return float(TAU_GRID[cell % 16]), float(L_GRID[cell // 16]), 20240 + p


def make_sweep(first, count, eta0):
    """fields eta0 * (1 + 0.05 z_p), z_p ~ N(0,1) per node from default_rng(20240 + p)"""
    taus, Ls = np.zeros(count), np.zeros(count)
    eta = np.zeros((count, len(eta0)))
    for i in range(count):
        tau, L, seed = sweep_params(first + i)
        taus[i], Ls[i] = tau, L
        eta[i] = eta0 * (1 + 0.05 * np.random.default_rng(seed).standard_normal(len(eta0)))
    return taus, Ls, eta


def shard(total, rank, world):
    """contiguous block [p0, p1) of rank `rank`: problem p -> rank p*world//total"""
    p0 = (rank * total + world - 1) // world
    p1 = ((rank + 1) * total + world - 1) // world
    return p0, p1


def gather_results(local, total, rank, world):
    """all-gather per-problem result rows (local: [p1-p0, k] float64) into [total, k] on every rank"""
    local = np.ascontiguousarray(local, dtype=np.float64)
    if world == 1:
        return local
    import torch
    import torch.distributed as dist
    k = local.shape[1]
    width = max(shard(total, r, world)[1] - shard(total, r, world)[0] for r in range(world))
    backend = dist.get_backend()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))) if backend == "nccl" else torch.device("cpu")
    buf = torch.zeros((width, k), dtype=torch.float64, device=dev)
    buf[: local.shape[0]] = torch.from_numpy(local).to(dev)
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    full = np.zeros((total, k))
    for r in range(world):
        p0, p1 = shard(total, r, world)
        full[p0:p1] = out[r][: p1 - p0].cpu().numpy()
    return full


def run_sharded(total, rank, world, evaluate):
    """evaluate(p0, p1) -> [p1-p0, k] results of this rank's block; returns the gathered [total, k]."""
    p0, p1 = shard(total, rank, world)
    return gather_results(evaluate(p0, p1), total, rank, world)


# ------------------------------------------------------------------------------------------------
# time-to-converge of sweep problems: the reference's continuation flow (drivescft.cc:259-322: solve on N=33,
# refine every cell, transfer the field by spline, solve again ... up to the target mesh) with the device-resident
# Broyden solver on every level.  One set of per-level engines is reused for all problems of a rank.
class Continuation:
    """levels N = 33, 65, ..., N_target on uniform meshes; scheme/nsteps as in Engine."""

    def __init__(self, N_target=1025, N0=33, nsteps=2048, scheme=None, device=0, tol=1e-9, retries=2):
        from . import engine as E
        self.E = E
        self.tol, self.retries = tol, retries
        self.levels = []
        N = N0
        while True:
            self.levels.append(N)
            if N >= N_target:
                break
            N = 2 * N - 1
        if self.levels[-1] != N_target:
            raise ValueError(f"N_target={N_target} is not reachable from N0={N0} by bisection")
        scheme = E.IE_ROWSCALE if scheme is None else scheme
        self.engines = [E.Engine(N, nsteps=nsteps, scheme=scheme, max_batch=N - 2, device=device) for N in self.levels]

    def close(self):
        for e in self.engines:
            e.close()
        self.engines = []

    def solve(self, tau, L, eta_mid0):
        """eta_mid0: interior field on the coarsest level.  Returns dict(check, err, eta_mid, F, Q, level_err)."""
        E = self.E
        eta = np.ascontiguousarray(eta_mid0, dtype=np.float64)
        assert len(eta) == self.levels[0] - 2
        level_err, check, err, attempts = [], 1, float("nan"), 0
        for lvl, (N, eng) in enumerate(zip(self.levels, self.engines)):
            eng.set_problem(-1, tau, L)
            if lvl:
                _, eta = E.refine_mesh(np.linspace(0.0, L, self.levels[lvl - 1]), eta)
            for attempt in range(1 + self.retries):
                # a line-search stall (check=1) just above the tolerance restarts from the returned field with a
                # fresh Jacobian, which is what a user of broydn.c does by calling it again (1D_FEM.c:356)
                _, check, eta, err, _ = eng.broydn_device(eta, self.tol, keep_trial=True)
                attempts += 1
                if check == 0 or not np.isfinite(err):
                    break
            level_err.append(err)
            if check != 0:
                break
        eng = self.engines[len(level_err) - 1]
        eng.residual(eta)  # phi, Q and the boundary values of the returned field for F
        return dict(check=check, err=err, eta_mid=eta, F=eng.free_energy(0, f0bar=0.0), Q=eng.Q(0), level_err=level_err, attempts=attempts,
                    N=self.levels[len(level_err) - 1])


def converge_block(p0, p1, eta33_mid, N_target=1025, nsteps=2048, scheme=None, device=0, tol=1e-9, threads=1):
    """Converge sweep problems [p0, p1) (sweep_params; initial field eta33_mid * (1 + 0.05 z_p) on N=33).
    Returns rows [check, err, F, Q, seconds].  threads > 1: that many host threads, each with its own set of
    per-level engines (own stream, own Broyden state), take problems from a shared queue, so that the serial dense
    algebra of one problem overlaps with the marches of the others on the same GPU."""
    import queue
    import time
    from concurrent.futures import ThreadPoolExecutor
    rows = np.zeros((p1 - p0, 5))
    threads = max(1, min(threads, p1 - p0))
    conts = [Continuation(N_target, len(eta33_mid) + 2, nsteps, scheme, device, tol) for _ in range(threads)]
    pool = queue.SimpleQueue()
    for c in conts:
        pool.put(c)

    def one(i):
        cont = pool.get()
        try:
            tau, L, seed = sweep_params(p0 + i)
            z = np.random.default_rng(seed).standard_normal(len(eta33_mid))
            t0 = time.perf_counter()
            r = cont.solve(tau, L, eta33_mid * (1 + 0.05 * z))
            rows[i] = (r["check"], r["err"], r["F"], r["Q"], time.perf_counter() - t0)
        finally:
            pool.put(cont)

    try:
        t0 = time.perf_counter()
        if threads == 1:
            for i in range(p1 - p0):
                one(i)
        else:
            with ThreadPoolExecutor(threads) as ex:
                list(ex.map(one, range(p1 - p0)))
        converge_block.last_wall = time.perf_counter() - t0   # solve phase only (engines already created)
    finally:
        for c in conts:
            c.close()
    return rows


converge_block.last_wall = 0.0


# ------------------------------------------------------------------------------------------------
# The same continuation for a whole block of the sweep in lock-step on the device (scftb_sweep_*): preconditioned
# Anderson mixing on every level, fields, history rings and the mesh transfer resident in HBM.
def converge_block_batched(p0, p1, eta33_mid, levels=6, nsteps=2048, scheme=None, device=0, tol=1e-9, nn=10, solver=None,
                           want_fields=False):
    """Converge sweep problems [p0, p1).  Returns dict(rows [p1-p0, 7] = (status, err, evaluations, Q, F,
    evaluations on the target mesh, N reached), seconds (solve phase), level_seconds, eta (if want_fields))."""
    import time
    from . import engine as E
    count = p1 - p0
    t_ms = time.perf_counter()
    taus, Ls, eta0 = make_sweep(p0, count, np.asarray(eta33_mid, dtype=np.float64))
    t_ms = time.perf_counter() - t_ms
    own = solver is None
    if own:
        solver = E.SweepSolver(count, N0=len(eta33_mid) + 2, levels=levels, nsteps=nsteps,
                               scheme=E.IE_ROWSCALE if scheme is None else scheme, tol=tol, nn=nn, device=device)
    try:
        t0 = time.perf_counter()
        r = solver.solve(taus, Ls, eta0, want_fields=want_fields)
        r["seconds"] = time.perf_counter() - t0
        r["seconds_make_sweep"] = t_ms
    finally:
        if own:
            solver.close()
    return r
