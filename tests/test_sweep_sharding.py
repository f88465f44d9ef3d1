"""N>1 path on CPU: world_size-2 gloo run of the sweep sharding + final all-gather.
The per-problem evaluation is replaced by a cheap deterministic function of the problem's
(tau, L, field) so that no GPU is needed; the GPU box runs the same code with NCCL (bench.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_blocks_partition_the_sweep():
    from scft_b200.sweep import shard
    for total in (1, 7, 4096, 4099):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_sweep_params_grid():
    from scft_b200.sweep import sweep_params, make_sweep
    cells = {sweep_params(p)[:2] for p in range(256)}
    assert len(cells) == 256
    assert sweep_params(0)[:2] == sweep_params(256)[:2] and sweep_params(0)[2] != sweep_params(256)[2]
    taus, Ls, eta = make_sweep(5, 3, np.ones(9))
    t2, L2, e2 = make_sweep(6, 1, np.ones(9))
    assert taus[1] == t2[0] and np.array_equal(eta[1], e2[0])


def _fake_eval(p0, p1):
    from scft_b200.sweep import make_sweep
    taus, Ls, eta = make_sweep(p0, p1 - p0, np.linspace(-1, 1, 31))
    return np.stack([taus * Ls, eta.sum(axis=1), np.arange(p0, p1, dtype=np.float64)], axis=1)


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from scft_b200.sweep import run_sharded
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = run_sharded(total, rank, world, _fake_eval)
    q.put((rank, full))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [64, 37])
def test_two_rank_gloo_sweep_equals_single_process(total):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _fake_eval(0, total)
    for r in range(2):
        assert np.array_equal(got[r], ref)


def test_continuation_level_ladder_is_checked_before_any_engine_is_created():
    """33 -> 65 -> ... by bisection only: an unreachable target is an argument error (no GPU needed to say so)"""
    import pytest
    from scft_b200 import sweep
    with pytest.raises(ValueError, match="not reachable"):
        sweep.Continuation(N_target=1000, N0=33)


def test_sweep_params_cover_the_tau_L_grid():
    from scft_b200 import sweep
    cells = {sweep.sweep_params(p)[:2] for p in range(256)}
    assert len(cells) == 256                                  # 16 x 16 distinct (tau, L) cells
    assert sweep.sweep_params(0)[:2] == sweep.sweep_params(256)[:2]   # then the seed axis
    assert sweep.sweep_params(0)[2] != sweep.sweep_params(256)[2]
