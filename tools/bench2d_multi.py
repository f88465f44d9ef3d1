"""2-D path on `world` GPUs (torchrun): slab partition + NCCL.  usage: torchrun ... tools/bench2d_multi.py nx ny nsteps"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import scft_b200 as sb

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nx, ny, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt = torch.tensor(list(sb.nccl_unique_id()), dtype=torch.uint8, device="cuda")
dist.broadcast(idt, 0)
fx = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/ref_fixtures.npz"))
L = sb.L_REF
x = L * np.arange(nx + 1) / nx
eta_x = np.interp(x, fx["res1024_xl"] * L, fx["res1024_eta"])
y = np.arange(ny + 1) / ny
eta = (eta_x[:, None] * (1 + 0.1 * np.cos(2 * np.pi * y)[None, :])).ravel()
eng = sb.Engine2D(nx, ny, L=L, Ly=L, nsteps=n, rtol=1e-12, device=local, rank=rank, world=world, nccl_id=bytes(idt.cpu().tolist()))
if len(sys.argv) > 4 and sys.argv[4] == "p2p":
    hb = torch.tensor(list(eng.p2p_handle()), dtype=torch.uint8, device="cuda")
    allh = [torch.zeros_like(hb) for _ in range(world)]
    dist.all_gather(allh, hb)
    eng.p2p_attach(b"".join(bytes(h.cpu().tolist()) for h in allh))
    dist.barrier()
for rep in range(2):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter(); out = eng.residual(eta); torch.cuda.synchronize(); wall = time.perf_counter() - t0
it, ms = eng.stats()
t = torch.tensor([ms], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
chk = torch.tensor([float(np.abs(out).sum())], device="cuda", dtype=torch.float64); dist.all_reduce(chk)
if rank == 0:
    ndof = (nx + 1) * (ny + 1)
    ms = float(t[0])
    print(f"world {world} [{sys.argv[4] if len(sys.argv) > 4 else 'nccl'}]: mesh {nx}x{ny} ({ndof} DOFs), {n} steps: march {ms:.1f} ms, {it} CG iterations = {it/n:.1f}/step, "
          f"{ms*1e3/it:.2f} us/iteration, {ndof*n/(ms*1e-3):.3e} DOF-steps/s, sum|out| {float(chk[0]):.12e}", flush=True)
if len(sys.argv) > 4 and sys.argv[4] == "p2p":
    eng.p2p_detach()
dist.barrier()
eng.close()
dist.destroy_process_group()
