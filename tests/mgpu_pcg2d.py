"""Slab-partitioned 2-D march over `world` GPUs (run under torchrun, one rank per GPU):
every rank's phi rows must equal the single-GPU result and the oracle.  Prints MGPU2D_OK on rank 0."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import scft_b200 as sb
    from oracle import oracle as O, oracle2d as O2
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    def fresh_id():
        # rank 0 creates the NCCL id of the engine's own communicator (one id per communicator);
        # the 128 bytes travel through torch.distributed
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(sb.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        return bytes(idt.cpu().tolist())

    mode = sys.argv[1] if len(sys.argv) > 1 else "nccl"   # "p2p": peer-memory persistent kernel instead of NCCL
    ok = True
    for (nx, ny, n) in [(31, 5, 16), (64, 16, 32), (255, 63, 16)]:
        L, Ly = O.L_REF, O.L_REF * ny / nx
        rng = np.random.default_rng(nx)
        x = L * np.arange(nx + 1) / nx
        eta = (3.0 * np.cos(2 * np.pi * x / L)[:, None] + rng.standard_normal((nx + 1, ny + 1))).ravel()
        eng = sb.Engine2D(nx, ny, L=L, Ly=Ly, nsteps=n, rtol=1e-13, device=local, rank=rank, world=world, nccl_id=fresh_id())
        if mode == "p2p":
            hb = torch.tensor(list(eng.p2p_handle()), dtype=torch.uint8, device="cuda")
            allh = [torch.zeros_like(hb) for _ in range(world)]
            dist.all_gather(allh, hb)
            eng.p2p_attach(b"".join(bytes(h.cpu().tolist()) for h in allh))
            dist.barrier()
        out = eng.residual(eta)
        ref = O2.residual(nx, ny, L, Ly, eta, nsteps=n)
        sl = slice(eng.row0, eng.row0 + eng.nrows)
        e1 = np.abs(eng.phi() - ref["phi"][sl]).max() / np.abs(ref["phi"]).max()
        e2 = np.abs(out - ref["out"][sl]).max()
        it, ms = eng.stats()
        print(f"rank {rank} [{mode}]: mesh {nx}x{ny} rows [{eng.row0},{eng.row0 + eng.nrows}) phi rel err {e1:.2e} out err {e2:.2e} "
              f"cg iterations {it} march {ms:.1f} ms", flush=True)
        ok = ok and e1 < 1e-10 and e2 < 1e-10
        if mode == "p2p":
            eng.p2p_detach()
        dist.barrier()
        eng.close()
        dist.barrier()
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MGPU2D_OK" if t.item() == 1.0 else "MGPU2D_FAIL", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
