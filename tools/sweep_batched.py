"""The 4096-problem sweep (BASELINE.json configs[2]) to convergence with the batched continuation solver.

  python tools/sweep_batched.py [problems_total=4096] [levels=6]
  python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/sweep_batched.py

Each rank converges its contiguous block of the sweep (sweep.shard) — N=33 -> ... -> 1025, n=2048, IE row-scaled,
tol 1e-9 on max|phi0 - phi| — and the per-problem rows are all-gathered.  No extrapolation: every problem is solved.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from scft_b200 import sweep, engine as E  # noqa: E402


def main():
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    levels = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl")
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ref_fixtures.npz"))
    eta33 = fx["n33_eta"][1:-1]
    p0, p1 = sweep.shard(total, rank, world)
    t0 = time.perf_counter()
    solver = E.SweepSolver(p1 - p0, N0=33, levels=levels, device=local)
    t_create = time.perf_counter() - t0
    first = None
    for rep in range(2):      # second pass: kernels loaded, per-cell free-energy weights cached in the solver
        r = sweep.converge_block_batched(p0, p1, eta33, levels=levels, device=local, solver=solver)
        first = r["seconds"] if first is None else first
    solver.close()
    rows = np.hstack([r["rows"], np.full((p1 - p0, 1), r["seconds"])])
    full = sweep.gather_results(rows, total, rank, world)
    if rank == 0:
        ok = full[:, 0] == 0
        secs = full[:, 7].max()
        N = (33 - 1) * 2 ** (levels - 1) + 1
        print(f"world {world}: {total} sweep problems to N={N}, n=2048, IE row-scaled, tol 1e-9: {int(ok.sum())} converged, "
              f"worst residual {np.nanmax(full[ok, 1]):.2e}; slowest rank {secs:.3f} s ({total / secs:.0f} problems/s whole job, "
              f"{total / world / secs:.0f} per GPU); first (cold) pass on rank 0 {first:.3f} s; solver setup {t_create:.2f} s")
        print(f"  evaluations per problem: all levels mean {full[:, 2].mean():.1f} max {full[:, 2].max():.0f}; target mesh mean "
              f"{full[:, 5].mean():.1f} max {full[:, 5].max():.0f}")
        print(f"  rank 0 seconds per level {np.array2string(r['level_seconds'], precision=4)}")
        print(f"  F range {np.nanmin(full[:, 4]):.6e} .. {np.nanmax(full[:, 4]):.6e}, Q range {np.nanmin(full[:, 3]):.6f} .. {np.nanmax(full[:, 3]):.6f}")
        if total % 256 == 0 and total >= 512:   # the seeds of one (tau, L) cell must end in the same state
            F = full[:, 4].reshape(-1, 256)
            Qs = full[:, 3].reshape(-1, 256)
            print(f"  seeds of a (tau, L) cell agree: max relative spread of the free energy {np.nanmax(np.ptp(F, axis=0) / np.abs(F[0])):.2e}, "
                  f"of Q {np.nanmax(np.ptp(Qs, axis=0) / np.abs(Qs[0])):.2e} (over {F.shape[0]} seeds x 256 cells)")
        for i in np.flatnonzero(~ok)[:8]:
            print(f"  not converged: problem {i} status {int(full[i, 0])} err {full[i, 1]:.2e} reached N={int(full[i, 6])}")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
