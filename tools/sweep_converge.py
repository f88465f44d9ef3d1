"""Time-to-converge of the 4096-problem sweep, measured on a bounded sample.

  python tools/sweep_converge.py [problems_per_rank=32] [N_target=1025] [host_threads=1]
  python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/sweep_converge.py 32

Each rank converges the first `problems_per_rank` problems of its block of the sweep (consecutive problems walk
through the (tau, L) grid cells) by the continuation flow N=33 -> ... -> N_target with the device-resident Broyden
solver, tolerance 1e-9 on max|phi0 - phi|.  The rows are all-gathered and rank 0 prints the statistics and the
explicit extrapolation to the whole sweep.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from scft_b200 import sweep  # noqa: E402


def main():
    count = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    N_target = int(sys.argv[2]) if len(sys.argv) > 2 else 1025
    threads = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl")
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ref_fixtures.npz"))
    eta33 = fx["n33_eta"][1:-1]  # DEALII_SCFT/inputFiles/N=33_for_read.txt (the reference's own start, drivescft.cc:264)
    total = 4096
    p0, p1 = sweep.shard(total, rank, world)
    rows = sweep.converge_block(p0, p0 + count, eta33, N_target=N_target, device=local, threads=threads)
    wall = sweep.converge_block.last_wall
    full = sweep.gather_results(np.hstack([rows, np.full((count, 1), wall)]), count * world, rank, world) \
        if world > 1 else np.hstack([rows, np.full((count, 1), wall)])
    if rank == 0:
        ok = full[:, 0] == 0
        sec = full[:, 4]
        print(f"world {world}: {len(full)} sweep problems to N={N_target}, n=2048, IE row-scaled, tol 1e-9: "
              f"{int(ok.sum())} converged, max err {np.nanmax(full[:, 1]):.2e}, "
              f"seconds per problem mean {sec.mean():.3f} min {sec.min():.3f} max {sec.max():.3f}; "
              f"slowest rank {full[:, 5].max():.1f} s for {count} problems")
        print(f"  F range {np.nanmin(full[:, 2]):.6e} .. {np.nanmax(full[:, 2]):.6e}, Q range "
              f"{np.nanmin(full[:, 3]):.6f} .. {np.nanmax(full[:, 3]):.6f}")
        per_rank = total // world
        rate = count / full[:, 5].max()
        print(f"  {threads} host thread(s) per GPU: {rate:.2f} problems/s per GPU; extrapolated to the whole sweep: "
              f"{per_rank} problems per GPU / {rate:.2f} per s = {per_rank / rate / 60:.1f} min on {world} GPU(s)")
        for i in np.flatnonzero(~ok)[:8]:
            print(f"  not converged: sample {i} check {int(full[i, 0])} err {full[i, 1]:.2e}")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
