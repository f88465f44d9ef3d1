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
