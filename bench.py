#!/usr/bin/env python
"""bench.py — propagator DOF-steps/s of the SCFT hot path on N B200s (BASELINE.json metric).

Workload (BASELINE.json configs[2]): a sweep of independent 1D hard-surface
SCFT problems over (tau, L, perturbed eta0), m=1024 elements (N=1025 nodes, 1023 unknowns),
n=2048 implicit-Euler contour steps, P=1 propagator sweep per evaluation (the reference's
symmetric one-sweep form q+(x,s)=q(x,1-s), drivescft.cc:189-190).  One "step" is one SCFT
iteration of every problem of the batch: a residual evaluation (the march kernel) followed by
the field update (preconditioned Anderson mixing, scft_b200/csrc/pmixer.cu) on the target mesh, started from the
fields the continuation hands to that mesh (coarser levels solved, field transferred by spline).  The 4096 problems of the sweep are sharded across the ranks in contiguous blocks with no
data-path collective: STRONG scaling (4096 / N problems per GPU), as BASELINE.json words configs[2].  At N > 1 a
second, weak-scaling leg (4096 problems per GPU) is reported under the extra key "weak".

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

One JSON line on stdout (rank 0).  Timing: CUDA events on the launching stream, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_NODES = 1025
NSTEPS = 2048
SCHEME = 0  # IE, row-scaled C: the scheme of 1D_FEM.c:95-186
TOTAL_PROBLEMS = 4096   # the sweep of BASELINE.json configs[2]; sharded over the ranks
BYTES_PER_DOF_STEP = 8.0  # lean history: each q(x_i,s_j), j<n/2, is written once and read once => 4+4 B per DOF-step


def make_sweep(first, count):
    """problems [first, first+count) of the sweep (scft_b200/sweep.py) with fields on the target
    mesh: the perturbed spectral guess (used by the CPU legs, which only evaluate residuals)"""
    from scft_b200 import sweep
    fx = np.load(os.path.join(ROOT, "tests", "golden", "ref_fixtures.npz"))
    return sweep.make_sweep(first, count, fx["res1024_eta"][1:-1])


def eta33_start():
    """the reference's own start field (DEALII_SCFT/inputFiles/N=33_for_read.txt, drivescft.cc:264), interior nodes"""
    fx = np.load(os.path.join(ROOT, "tests", "golden", "ref_fixtures.npz"))
    return fx["n33_eta"][1:-1]


LEVELS = 6   # 33 -> 65 -> 129 -> 257 -> 513 -> 1025 (drivescft.cc:291-322: refine every cell, solve again)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self, t_begin=None, t_end=None):
        """median SM clock over the samples taken inside [t_begin, t_end] (host wall clock)"""
        import datetime
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        total = 0
        for ln in out.strip().splitlines():
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 8:
                continue
            total += 1
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if t_begin is not None and not (t_begin - 0.02 <= ts <= t_end + 0.02):
                    continue
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples_in_timed_region": len(sm), "samples": total,
                "reasons": sorted(reasons)}


_emit = None


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def fair_cpu_throughput(first, count, threads):
    """DOF-steps/s of the fair CPU port (oracle/scft_fast.c: Thomas once per field, half history, 8 problems per SIMD
    group, `threads` POSIX threads) on problems [first, first+count) of the sweep."""
    from oracle import oracle as O
    O.lib()
    taus, Ls, eta = make_sweep(first, count)
    t0 = time.perf_counter()
    r = O.fast_sweep(taus, Ls, eta, N_NODES, scheme=SCHEME, nsteps=NSTEPS, threads=threads, want_phi=False)
    dt = time.perf_counter() - t0
    assert np.isfinite(r["out"]).all()
    return count * (N_NODES - 2) * NSTEPS / dt, dt


def checker_throughput(count, threads):
    """DOF-steps/s of the checker oracle (oracle/scft_oracle.c: pivoting band LU, full history) — round 1's CPU arm"""
    from oracle import oracle as O
    O.lib()
    taus, Ls, eta = make_sweep(0, count)

    def one(i):
        x = O.mesh_uniform(N_NODES, Ls[i])
        return O.residual(O.eta_full(x, eta[i]), O.f0_given(x, taus[i]), scheme=SCHEME, nsteps=NSTEPS, L=Ls[i])["Q"]

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(one, range(count)))
    dt = time.perf_counter() - t0
    return count * (N_NODES - 2) * NSTEPS / dt, dt


def reference_faithful_leg(evals=3):
    """The reference CPU path as the reference runs it (SURVEY.md 8d "reference-faithful, 1 thread"):
    ONE host thread, the reference's own adm_chen (ADM_chen_C.c, compiled in place into oracle/_ref) driving a residual
    whose field extension is the reference's own spline_chen + dense gaussj (spline_chen.c:23-68 from scft.cc:474,
    O(N^3) per evaluation) followed by the oracle's IRK4 march and romint quadrature (deal.II / UMFPACK themselves are
    unbuildable here).  m = 1024: a few evaluations only, stated."""
    from oracle import oracle as O
    if not O.have_ref():
        return None
    x = O.mesh_uniform(N_NODES)
    f0 = O.f0_given(x)
    eta0 = make_sweep(0, 1)[2][0]
    calls = []

    def F(em):
        t0 = time.perf_counter()
        ef = O.ref_spline(x[1:-1], em, x)
        t1 = time.perf_counter()
        out = O.residual(ef, f0, scheme=O.IRK4_CONSISTENT, nsteps=NSTEPS)["out"]
        calls.append((t1 - t0, time.perf_counter() - t1))
        return out

    t0 = time.perf_counter()
    O.ref_adm_chen(F, eta0, 1e-30, evals - 1, 0.99, 2)
    dt = time.perf_counter() - t0
    n = len(calls)
    return {"what": "reference adm_chen (oracle/_ref) driving reference spline_chen+gaussj + oracle IRK4 march + romint, "
                    "m=1024 n=2048, 1 host thread", "evaluations": n, "seconds": dt,
            "seconds_per_evaluation": dt / n, "spline_gaussj_seconds_per_evaluation": sum(c[0] for c in calls) / n,
            "march_seconds_per_evaluation": sum(c[1] for c in calls) / n,
            "value": n * (N_NODES - 2) * NSTEPS / dt, "unit": "DOF-steps/s", "cores": 1}


def run_reference(args, rank):
    if rank != 0:
        return
    cores = host_cores()
    per_step = 512                      # one eighth of the sweep per step: a fraction of a second on 16 threads
    for _ in range(args.warmup):
        fair_cpu_throughput(0, 64, cores)
    t = []
    for k in range(args.steps):
        v, dt = fair_cpu_throughput((k * per_step) % TOTAL_PROBLEMS, per_step, cores)
        t.append(dt)
    tot = sum(t)
    value = args.steps * per_step * (N_NODES - 2) * NSTEPS / tot
    sample = (f"{per_step} problems of the 4096-problem sweep per step, fair CPU port oracle/scft_fast.c "
              f"(Thomas once per field, half history, AVX SIMD over 8 problems), {cores} host threads")
    line = {"impl": "reference", "metric": "propagator_dof_steps_per_s", "value": value, "unit": "DOF-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(-(-TOTAL_PROBLEMS // max(args.gpus, 1)), max(args.gpus, 1)),
            "scft_residual_evaluations_per_s": args.steps * per_step / tot,
            "cpu_baseline": {"value": value, "unit": "DOF-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "DOF-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


def workload_config(problems_per_gpu, world, total=TOTAL_PROBLEMS):
    return {"workload": "sweep of 4096 independent 1D hard-surface SCFT problems (tau x L x eta0-seed grid), "
                        "m=1024 n=2048 implicit Euler (BASELINE.json configs[2])",
            "problems_total": total, "problems_per_gpu": problems_per_gpu, "N": N_NODES, "unknowns": N_NODES - 2,
            "nsteps": NSTEPS, "scheme": "IE_ROWSCALE (1D_FEM.c:95-186)", "propagator_sweeps_P": 1,
            "step": "one SCFT iteration of every problem: residual evaluation + Anderson field update",
            "skipped_problems": 0,
            "l2": "inputs larger than L2: each step streams the q history (> 0.5 GB per GPU even at N=8) through HBM",
            "parallelism": f"4096 problems sharded over {world} rank(s) in contiguous blocks, no data-path collective"}


class SweepRun:
    """problems [p0, p1) of the sweep resident on this rank's GPU: engine + device-resident preconditioned mixer.
    The start fields of the timed iterations are what the continuation hands to the target mesh: levels 33..513 solved
    (untimed preparation here; the whole flow is timed by the sweep_converged leg), field cut to 1025 nodes by spline."""

    def __init__(self, p0, p1, local, torch, scft_b200):
        import ctypes as C
        from scft_b200 import sweep
        self.P = p1 - p0
        dev = torch.device("cuda", local)
        self.taus, self.Ls, eta0 = sweep.make_sweep(p0, self.P, eta33_start())
        coarse = scft_b200.SweepSolver(self.P, N0=33, levels=LEVELS - 1, nsteps=NSTEPS, scheme=SCHEME, device=local)
        r = coarse.solve(self.taus, self.Ls, eta0)
        coarse.close()
        self.coarse_converged = int((r["rows"][:, 0] == 0).sum())
        d_c = torch.from_numpy(r["eta"]).to(dev)
        d_L = torch.from_numpy(self.Ls).to(dev)
        self.d_eta = torch.zeros((self.P, N_NODES - 2), dtype=torch.float64, device=dev)
        rc = scft_b200.lib().scftb_refine_uniform_batch_device(self.P, (N_NODES + 1) // 2, C.c_void_p(d_L.data_ptr()),
                                                               C.c_void_p(d_c.data_ptr()), C.c_void_p(self.d_eta.data_ptr()), None)
        assert rc == 0
        self.eng = scft_b200.Engine(N_NODES, nsteps=NSTEPS, scheme=SCHEME, max_batch=self.P, device=local)
        for p in range(self.P):
            self.eng.set_problem(p, self.taus[p], self.Ls[p])
        self.h_eta = self.d_eta.cpu().pin_memory()
        self.eta = self.h_eta.numpy()
        self.h_out = torch.empty_like(self.h_eta).pin_memory()
        self.stream = torch.cuda.current_stream()
        # tol = 0: no problem is ever frozen, so EVERY problem is evaluated and updated in EVERY step; the iterates reach
        # max|phi0 - phi| < 1e-9 after ~8 steps and stay at round-off level afterwards
        self.mixer = scft_b200.PrecondAndersonBatch(self.eng, self.P, tol=0.0, nn=10)
        self.mixer.reset_device(self.d_eta.data_ptr(), self.stream.cuda_stream)

    def step(self):
        """one SCFT iteration of the resident problems, fields resident in HBM"""
        self.mixer.iterate_device(self.stream.cuda_stream)

    def reset(self):
        self.mixer.reset_device(self.d_eta.data_ptr(), self.stream.cuda_stream)

    def close(self):
        self.mixer.close()
        self.eng.close()


def timed_steps(run, steps, warmup, barrier, torch):
    """(ms for `steps` steps on this rank, mean march-kernel ms, launches): CUDA events on the launching stream"""
    import scft_b200
    for _ in range(warmup):
        run.step()
    run.reset()                       # timed iterations start from the sweep's own fields
    run.eng.set_timing(True)
    barrier()
    scft_b200.launch_count(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    run.eng.march_ms()                # drop earlier launches
    ev0.record(run.stream)
    for _ in range(steps):
        run.step()
    ev1.record(run.stream)
    barrier()
    launches = scft_b200.launch_count()
    tot, cnt = run.eng.march_ms()
    run.eng.set_timing(False)
    return ev0.elapsed_time(ev1), tot / max(cnt, 1), launches


def parity_spot_check(run, count=8):
    """`count` problems of this rank's shard against the CPU oracle (checker): phi, Q, residual of the fields the e2e
    leg just evaluated through scftb_residual_batch.  Returns (checked, max relative error)."""
    from oracle import oracle as O
    O.lib()
    P = run.P
    slots = run.eng.slots()
    pts = sorted({0, P - 1, P // 2, min(P - 1, slots - 1), min(P - 1, slots), min(P - 1, 2 * slots), P // 3, (2 * P) // 3})[:count]
    out = run.h_out.numpy()
    worst = 0.0
    for p in pts:
        x = O.mesh_uniform(N_NODES, run.Ls[p])
        ref = O.residual(O.eta_full(x, run.eta[p]), O.f0_given(x, run.taus[p]), scheme=SCHEME, nsteps=NSTEPS, L=run.Ls[p])
        scale = np.abs(ref["phi"]).max()
        worst = max(worst, np.abs(run.eng.phi(p) - ref["phi"]).max() / scale, abs(run.eng.Q(p) - ref["Q"]) / abs(ref["Q"]),
                    np.abs(out[p] - ref["out"]).max() / scale)
    return len(pts), float(worst)


def mesh2d_leg(rank, world, local, torch, dist, scft_b200):
    """Extra key "mesh2d": the 2-D path (BASELINE.json configs[3], [4]; pcg2d.cu) — Q1 mesh, matrix-free rows of A + ds(B + C),
    Jacobi-PCG per contour step, implicit Euler.
      * 1M DOFs (nx=1024, ny=1023 cells), n=2048: one full residual evaluation of a y-modulated field, sharded in x-slabs
        over the ranks (peer-memory persistent kernel at N > 1), CUDA-event time of the march, max over ranks.
      * N = 1 only: the same mesh with a y-invariant field against the 1-D engine (parity at configuration scale).
      * 16.8M DOFs (4095 x 4095 cells): microseconds per CG iteration with the iteration count per step capped
        (16 steps x 50 iterations: the per-iteration work is that of the real solve; the fields are not used)."""
    dev = torch.device("cuda", local)
    fx = np.load(os.path.join(ROOT, "tests", "golden", "ref_fixtures.npz"))
    L = scft_b200.L_REF

    def nccl_id():
        if world == 1:
            return None
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt = torch.tensor(list(scft_b200.nccl_unique_id()), dtype=torch.uint8, device=dev)
        dist.broadcast(idt, 0)
        return bytes(idt.cpu().tolist())

    def attach(eng):
        hb = torch.tensor(list(eng.p2p_handle()), dtype=torch.uint8, device=dev)
        allh = [torch.zeros_like(hb) for _ in range(world)]
        dist.all_gather(allh, hb)
        eng.p2p_attach(b"".join(bytes(h.cpu().tolist()) for h in allh))
        dist.barrier()

    def run(nx, ny, n, eta, maxit=0, mode="p2p", reps=2):
        eng = scft_b200.Engine2D(nx, ny, L=L, Ly=L * ny / nx, nsteps=n, rtol=1e-12, maxit=maxit, device=local, rank=rank,
                                 world=world, nccl_id=nccl_id())
        if world > 1 and mode == "p2p":
            attach(eng)
        for _ in range(reps):
            if world > 1:
                dist.barrier()
            out = eng.residual(eta)
        it, ms = eng.stats()
        phi = eng.phi()
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if mode == "p2p":
                eng.p2p_detach()
            dist.barrier()
        rows = (eng.row0, eng.nrows)
        eng.close()
        if world > 1:
            dist.barrier()
        return float(t[0]), it, phi, out, rows

    res = {"scheme": "implicit Euler on the deal.II matrices A, B, C (scft.cc:643-656), Jacobi-PCG per step, rtol 1e-12",
           "bytes_per_dof_per_cg_iteration": 128,
           "matrix": "matrix-free Q1 rows (closed-form element entries, eta read directly); 32 B SpMV pass + 96 B update pass"}
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(peaks_path))["hbm_gbs"] if os.path.exists(peaks_path) else 6650.0
    # ---- 1M DOFs, n = 2048, full evaluation
    nx, ny, n = 1024, 1023, NSTEPS
    x = L * np.arange(nx + 1) / nx
    eta_x = np.interp(x, fx["res1024_xl"] * L, fx["res1024_eta"])
    y = np.arange(ny + 1) / ny
    eta = (eta_x[:, None] * (1 + 0.1 * np.cos(2 * np.pi * y)[None, :])).ravel()
    ms, it, phi, out, rows = run(nx, ny, n, eta, reps=1 if world == 1 else 2)
    ndof = (nx + 1) * (ny + 1)
    res["dofs_1m"] = {"mesh_cells": [nx, ny], "dofs": ndof, "nsteps": n, "march_ms": ms, "cg_iterations": it,
                      "cg_iterations_per_step": it / n, "us_per_cg_iteration": ms * 1e3 / it,
                      "dof_steps_per_s": ndof * n / (ms * 1e-3),
                      "hbm_frac_per_gpu": 128.0 * ndof / world * it / (ms * 1e-3) / 1e9 / peak,
                      "exchange": "none (one GPU)" if world == 1 else "peer-memory persistent kernel (halo stores + in-kernel all-reduce over NVLink)"}
    if world == 1:
        # parity at configuration scale: y-invariant field == 1-D engine
        e1 = scft_b200.Engine(nx + 1, nsteps=n, scheme=scft_b200.IE_CONSISTENT, device=local)
        em = np.interp(x, fx["res1024_xl"] * L, fx["res1024_eta"])[1:-1]
        e1.residual(em)
        phi1, ef = e1.phi(), e1.eta_full()
        e1.close()
        ms2, it2, phi2, _, _ = run(nx, ny, n, np.repeat(ef, ny + 1), reps=1)
        phi2 = phi2.reshape(nx + 1, ny + 1)
        res["dofs_1m"]["parity_y_invariant_vs_1d_engine"] = {
            "max_abs_err_phi": float(np.abs(phi2 - phi1[:, None]).max()), "tolerance": 1e-10,
            "cg_iterations": it2, "march_ms": ms2}
    # ---- 16.8M DOFs, per-iteration cost
    nx = ny = 4095
    x = L * np.arange(nx + 1) / nx
    eta_x = np.interp(x, fx["res1024_xl"] * L, fx["res1024_eta"])
    y = np.arange(ny + 1) / ny
    eta = (eta_x[:, None] * (1 + 0.1 * np.cos(2 * np.pi * y)[None, :])).ravel()
    ndof = (nx + 1) * (ny + 1)
    big = {"mesh_cells": [nx, ny], "dofs": ndof, "nsteps": 16, "cg_iterations_per_step_cap": 50}
    for mode in (["one_gpu"] if world == 1 else ["p2p", "nccl"]):
        ms, it, _, _, _ = run(nx, ny, 16, eta, maxit=50, mode=mode, reps=2)
        big[mode] = {"march_ms": ms, "cg_iterations": it, "us_per_cg_iteration": ms * 1e3 / it,
                     "hbm_frac_per_gpu": 128.0 * ndof / world / (ms * 1e-3 / it) / 1e9 / peak}
    res["dofs_16m"] = big
    return res


def main():
    # stdout carries exactly ONE line (the JSON); anything libraries print (e.g. NCCL's version banner) goes to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    global _emit
    _emit = lambda line: (real_stdout.write(json.dumps(line) + "\n"), real_stdout.flush())
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--problems", type=int, default=TOTAL_PROBLEMS, help="problems of the whole sweep (sharded over the ranks)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="skip the extra weak-scaling leg at N > 1")
    ap.add_argument("--no-mesh2d", action="store_true", help="skip the extra 2-D mesh leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import scft_b200
    from scft_b200 import sweep

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    total = args.problems
    ni = N_NODES - 2
    p0, p1 = sweep.shard(total, rank, world)
    run = SweepRun(p0, p1, local, torch, scft_b200)
    P = run.P

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    # ---- device-resident timing (strong scaling: this rank's block of the 4096-problem sweep)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t_begin = time.time()
    ms, march_ms, launches = timed_steps(run, args.steps, args.warmup, barrier, torch)
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    ms_total, march_ms_avg = max_over_ranks([ms, march_ms])
    done, iters, err = run.mixer.status(run.stream.cuda_stream)
    kernel_name = run.eng.kernel_name()

    # ---- end to end through the C ABI with host buffers (H2D + kernel + D2H inside the call)
    barrier()
    for _ in range(2):
        run.eng.residual_host_ptr(P, run.h_eta.data_ptr(), run.h_out.data_ptr())
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run.eng.residual_host_ptr(P, run.h_eta.data_ptr(), run.h_out.data_ptr())
    torch.cuda.synchronize()
    e2e_s = max_over_ranks([time.perf_counter() - t0])[0]

    # ---- parity: a handful of this rank's problems against the CPU oracle (rank 0 reports; every rank checks)
    checked, max_rel = parity_spot_check(run)
    max_rel_all = max_over_ranks([max_rel])[0]

    # per-problem scalars of the whole sweep: the one collective of the workload, outside the timed region
    local_rows = np.stack([err, iters.astype(np.float64)], axis=1)
    allres = sweep.gather_results(local_rows, total, rank, world) if world > 1 else local_rows
    run.close()

    # ---- the sweep to convergence, for real: every problem of this rank's block from the N=33 start field through all
    # levels to max|phi0 - phi| < 1e-9 on the target mesh; wall clock from host start fields to host result rows
    solver = scft_b200.SweepSolver(P, N0=33, levels=LEVELS, nsteps=NSTEPS, scheme=SCHEME, tol=1e-9, device=local)
    conv = sweep.converge_block_batched(p0, p1, eta33_start(), levels=LEVELS, solver=solver)          # first pass (cold)
    conv_first_s = max_over_ranks([conv["seconds"]])[0]
    conv_passes = []
    for _ in range(3):   # three timed passes; the median is reported
        barrier()
        t0 = time.perf_counter()
        conv = sweep.converge_block_batched(p0, p1, eta33_start(), levels=LEVELS, solver=solver)
        torch.cuda.synchronize()
        conv_passes.append(max_over_ranks([time.perf_counter() - t0])[0])
    conv_s = float(np.median(conv_passes))
    solver.close()
    conv_rows = sweep.gather_results(conv["rows"], total, rank, world) if world > 1 else conv["rows"]

    # ---- extra: weak scaling (4096 problems on every GPU), N > 1 only
    weak = None
    if world > 1 and not args.no_weak:
        wrun = SweepRun(rank * TOTAL_PROBLEMS, (rank + 1) * TOTAL_PROBLEMS, local, torch, scft_b200)
        wms, wmarch, _ = timed_steps(wrun, args.steps, args.warmup, barrier, torch)
        wms, wmarch = max_over_ranks([wms, wmarch])
        wrun.close()
        weak = {"scaling": "weak", "problems_per_gpu": TOTAL_PROBLEMS, "problems_total": world * TOTAL_PROBLEMS,
                "value": world * TOTAL_PROBLEMS * ni * NSTEPS * args.steps / (wms * 1e-3), "unit": "DOF-steps/s",
                "ms_per_step": wms / args.steps, "march_kernel_ms": wmarch,
                "hbm_frac_per_gpu": TOTAL_PROBLEMS * ni * NSTEPS * BYTES_PER_DOF_STEP / (wmarch * 1e-3) / 1e9}

    # ---- extra: the same residual evaluations with the reference driver's own stepper (IRK4, scft.cc:671-693): device-resident
    # evaluations of this rank's block, CUDA events on the launching stream, max over ranks
    irk4 = None
    try:
        e4 = scft_b200.Engine(N_NODES, nsteps=NSTEPS, scheme=scft_b200.IRK4_CONSISTENT, max_batch=P, device=local)
        taus4, Ls4, _ = sweep.make_sweep(p0, P, eta33_start())
        for p in range(P):
            e4.set_problem(p, taus4[p], Ls4[p])
        d_eta4 = torch.from_numpy(make_sweep(p0, P)[2]).to(dev)
        d_out4 = torch.empty_like(d_eta4)
        st4 = torch.cuda.current_stream()
        for _ in range(2):
            e4.residual_device(P, d_eta4.data_ptr(), d_out4.data_ptr(), st4.cuda_stream)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(st4)
        for _ in range(5):
            e4.residual_device(P, d_eta4.data_ptr(), d_out4.data_ptr(), st4.cuda_stream)
        ev1.record(st4)
        barrier()
        ms4 = max_over_ranks([ev0.elapsed_time(ev1) / 5])[0]
        ok4 = bool(torch.isfinite(d_out4).all().item())
        irk4 = {"kernel": e4.kernel_name(), "problems_total": total, "ms_per_evaluation_of_all_problems": ms4,
                "dof_steps_per_s": total * ni * NSTEPS / (ms4 * 1e-3), "finite": ok4,
                "scheme": "IRK4_CONSISTENT (2-stage Gauss-Legendre as one complex tridiagonal solve per step)"}
        e4.close()
        del d_eta4, d_out4
    except Exception as exc:   # noqa: BLE001
        irk4 = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- extra: the 2-D mesh path (configs[3], [4]); a failure here must not take the headline line down
    mesh2d = None
    if not args.no_mesh2d:
        try:
            mesh2d = mesh2d_leg(rank, world, local, torch, dist, scft_b200)
        except Exception as exc:   # noqa: BLE001
            mesh2d = {"error": f"{type(exc).__name__}: {exc}"}

    if rank == 0:
        assert allres.shape[0] == total
        finite = int(np.isfinite(allres[:, 0]).sum())
        dof_steps_per_step = total * ni * NSTEPS
        value = dof_steps_per_step * args.steps / (ms_total * 1e-3)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        if weak:
            weak["hbm_frac_per_gpu"] /= peak
        Pmax = -(-total // world)   # the largest block: the rank whose kernel time is reported
        algo_bytes = Pmax * ni * NSTEPS * BYTES_PER_DOF_STEP
        traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum of one march launch (ncu --set full capture)
        for name in ("r2_traffic.json", "r1_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", name)
            if os.path.exists(tpath):
                tr = json.load(open(tpath))["march_kernel"]
                if tr["N"] == N_NODES and tr["nsteps"] == NSTEPS:   # per-launch traffic scales with the problem count
                    traffic = (tr["dram_bytes_read"] + tr["dram_bytes_write"]) * Pmax / tr["problems"]
                    break
        achieved = algo_bytes / (march_ms_avg * 1e-3) / 1e9
        line = {"metric": "propagator_dof_steps_per_s", "value": value, "unit": "DOF-steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(Pmax, world, total),
                "scft_iterations_per_s": total * args.steps / (ms_total * 1e-3),
                "problems_with_finite_residual_at_end": finite,
                "problems_below_1e-9_at_end": int((allres[:, 0] < 1e-9).sum()),
                "worst_residual_at_end": float(np.nanmax(allres[:, 0])),
                "field_update": "preconditioned Anderson mixing (pmix_kernel), window 10, never frozen (tol 0)",
                "sweep_converged": {
                    "problems": total, "converged": int((conv_rows[:, 0] == 0).sum()), "tol": 1e-9,
                    "seconds": conv_s, "problems_per_s": total / conv_s,
                    "seconds_all_passes": [round(v, 4) for v in conv_passes],   # "seconds" is their median
                    "seconds_first_pass": conv_first_s,   # cold: first launches of every level's kernels, per-cell free-energy weights not cached yet
                    "worst_residual": float(np.nanmax(conv_rows[:, 1])),
                    "evaluations_per_problem_mean": float(conv_rows[:, 2].mean()),
                    "evaluations_per_problem_max": float(conv_rows[:, 2].max()),
                    "target_mesh_evaluations_mean": float(conv_rows[:, 5].mean()),
                    "target_mesh_evaluations_max": float(conv_rows[:, 5].max()),
                    "free_energy_range": [float(np.nanmin(conv_rows[:, 4])), float(np.nanmax(conv_rows[:, 4]))],
                    "rank0_seconds_per_level_then_host_wait": [round(float(v), 4) for v in conv["level_seconds"]],
                    "rank0_seconds_start_fields": round(float(conv.get("seconds_make_sweep", 0.0)), 4),
                    "flow": "continuation N=33->65->129->257->513->1025 (drivescft.cc:291-322), preconditioned Anderson mixing on "
                            "every level, all problems of a rank in lock-step on the device; wall clock from host start fields "
                            "to host result rows, max over ranks, median of three passes after a cold one on the same solver object; no extrapolation"},
                "parity_checked": checked * world, "max_rel_err": max_rel_all,
                "parity": "phi, Q, residual of sampled problems vs the CPU oracle (oracle/scft_oracle.c), tolerance 1e-10",
                "clocks": clocks,
                "e2e": {"value": dof_steps_per_step * args.steps / e2e_s, "unit": "DOF-steps/s",
                        "h2d_bytes_per_step": P * ni * 8, "d2h_bytes_per_step": P * ni * 8,
                        "call": "scftb_residual_batch (pinned host buffers), wall clock, max over ranks; bytes per rank"},
                "gpu_launches": launches,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "kernel": kernel_name,
                             "problems_per_launch": Pmax,
                             "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": march_ms_avg,
                             "peak_source": peak_src,
                             "bytes_per_dof_step": BYTES_PER_DOF_STEP}}
        if weak:
            line["weak"] = weak
        if irk4:
            if "dof_steps_per_s" in irk4:
                irk4["hbm_frac_per_gpu"] = irk4["dof_steps_per_s"] / world * BYTES_PER_DOF_STEP / 1e9 / peak
            line["irk4"] = irk4
        if mesh2d:
            line["mesh2d"] = mesh2d
        assert max_rel_all < 1e-10, f"GPU results differ from the oracle: {max_rel_all:.3e}"
        if not args.no_cpu_baseline and world == 1:
            cores = host_cores()
            v, dt = fair_cpu_throughput(0, total, cores)   # the WHOLE workload once: a few seconds on 16 threads
            line["cpu_baseline"] = {"value": v, "unit": "DOF-steps/s", "cores": cores, "kind": "port",
                                    "sample": f"all {total} problems of the sweep once, fair CPU port oracle/scft_fast.c "
                                              f"(Thomas once per field, half history, SIMD over 8 problems), {dt:.1f} s"}
            cnt = 4 * max(cores, 8)
            cv, cdt = checker_throughput(cnt, cores)
            line["cpu_baseline"]["checker_port"] = {"value": cv, "unit": "DOF-steps/s", "cores": cores,
                                                    "sample": f"first {cnt} problems, oracle/scft_oracle.c (pivoting band LU, "
                                                              f"full history; round 1's CPU arm), {cdt:.1f} s"}
            rf = reference_faithful_leg()
            if rf:
                line["cpu_baseline"]["reference_faithful"] = rf
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
