#!/usr/bin/env python
"""bench.py — propagator DOF-steps/s of the SCFT hot path on N B200s (BASELINE.json metric).

Workload (BASELINE.json configs[2]): a sweep of independent 1D hard-surface
SCFT problems over (tau, L, perturbed eta0), m=1024 elements (N=1025 nodes, 1023 unknowns),
n=2048 implicit-Euler contour steps, P=1 propagator sweep per evaluation (the reference's
symmetric one-sweep form q+(x,s)=q(x,1-s), drivescft.cc:189-190).  One "step" is one SCFT
iteration of every problem of the batch: a residual evaluation (the march kernel) followed by
the field update.  Problems are sharded across ranks with no data-path collective (weak scaling:
4096 problems per GPU).

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
PROBLEMS_PER_GPU = 4096
BYTES_PER_DOF_STEP = 8.0  # lean history: each q(x_i,s_j), j<n/2, is written once and read once => 4+4 B per DOF-step


def make_sweep(first, count):
    """problems [first, first+count) of the sweep (scft_b200/sweep.py)"""
    from scft_b200 import sweep
    fx = np.load(os.path.join(ROOT, "tests", "golden", "ref_fixtures.npz"))
    return sweep.make_sweep(first, count, fx["res1024_eta"][1:-1])


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


def oracle_throughput(count, threads):
    """DOF-steps/s of the CPU oracle (the reference algorithm restated in C) on `count` problems
    of the same sweep, `threads` host threads (ctypes releases the GIL)."""
    from oracle import oracle as O
    O.lib()
    taus, Ls, eta = make_sweep(0, count)

    def one(i):
        x = O.mesh_uniform(N_NODES, Ls[i])
        f0 = O.f0_given(x, taus[i])
        ef = O.eta_full(x, eta[i])
        return O.residual(ef, f0, scheme=SCHEME, nsteps=NSTEPS, L=Ls[i])["Q"]

    t0 = time.perf_counter()
    if threads == 1:
        for i in range(count):
            one(i)
    else:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(one, range(count)))
    dt = time.perf_counter() - t0
    return count * (N_NODES - 2) * NSTEPS / dt, dt


def reference_code_probe():
    """One call of the reference's OWN spline_chen (dense gaussj on the (N-2)^2 spline matrix, spline_chen.c:23-68 called
    from scft.cc:474 on every residual evaluation of the deal.II flow) at N=1025, compiled in place into oracle/_ref.
    Informational: the step the unbuildable deal.II driver spends most of an m=1024 evaluation in."""
    from oracle import oracle as O
    if not O.have_ref():
        return None
    x = O.mesh_uniform(N_NODES)
    eta = make_sweep(0, 1)[2][0]
    t0 = time.perf_counter()
    O.ref_spline(x[1:-1], eta, x)
    dt = time.perf_counter() - t0
    return {"what": "reference spline_chen + gaussj, one residual evaluation's field extension at N=1025, 1 thread",
            "seconds": dt, "dof_steps_per_s_upper_bound": (N_NODES - 2) * NSTEPS / dt}


def run_reference(args, rank):
    if rank != 0:
        return
    cores = host_cores()
    per_step = max(cores, 8) * 16
    for _ in range(args.warmup):
        oracle_throughput(max(cores, 8), cores)
    t = []
    for _ in range(args.steps):
        v, dt = oracle_throughput(per_step, cores)
        t.append(dt)
    tot = sum(t)
    value = args.steps * per_step * (N_NODES - 2) * NSTEPS / tot
    sample = f"{per_step} problems of the sweep per step (of {PROBLEMS_PER_GPU} per GPU), {cores} host threads"
    line = {"impl": "reference", "metric": "propagator_dof_steps_per_s", "value": value, "unit": "DOF-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(per_step),
            "scft_residual_evaluations_per_s": args.steps * per_step / tot,
            "cpu_baseline": {"value": value, "unit": "DOF-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "DOF-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


def workload_config(problems_per_gpu):
    return {"workload": "sweep of independent 1D hard-surface SCFT problems (tau x L x eta0-seed grid), "
                        "m=1024 n=2048 implicit Euler (BASELINE.json configs[2])",
            "problems_per_gpu": problems_per_gpu, "N": N_NODES, "unknowns": N_NODES - 2, "nsteps": NSTEPS,
            "scheme": "IE_ROWSCALE (1D_FEM.c:95-186)", "propagator_sweeps_P": 1,
            "step": "one SCFT iteration of every problem: residual evaluation + Anderson field update",
            "skipped_problems": 0,
            "l2": "inputs larger than L2: each step streams the q history (>3 GB per GPU) through HBM",
            "parallelism": "problems sharded by rank, no data-path collective"}


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
    ap.add_argument("--problems", type=int, default=PROBLEMS_PER_GPU, help="problems per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    P = args.problems
    ni = N_NODES - 2

    taus, Ls, eta = make_sweep(rank * P, P)
    eng = scft_b200.Engine(N_NODES, nsteps=NSTEPS, scheme=SCHEME, max_batch=P, device=local)
    for p in range(P):
        eng.set_problem(p, taus[p], Ls[p])
    h_eta = torch.from_numpy(eta).pin_memory()
    h_out = torch.empty_like(h_eta).pin_memory()
    d_eta = h_eta.to(dev, non_blocking=True)
    d_out = torch.empty_like(d_eta)
    stream = torch.cuda.current_stream()
    # fixed-iteration benchmark: tol unreachable and freeze off, so EVERY problem is evaluated in EVERY step
    # (Anderson from the perturbed spectral guess diverges for a few percent of the problems; the reference
    # would exit(1) on their NaNs — here they keep costing the full march)
    mixer = scft_b200.AndersonBatch(eng, P, tol=1e-30, lmd=0.99, nn=2)
    mixer.set_freeze(False)
    mixer.reset_device(d_eta.data_ptr(), stream.cuda_stream)
    eng.set_timing(True)

    def step():
        """one SCFT iteration of the batch, fields resident in HBM"""
        mixer.iterate_device(stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    mixer.reset_device(d_eta.data_ptr(), stream.cuda_stream)   # timed iterations start from the sweep's fields
    barrier()
    scft_b200.launch_count(reset=True)
    t_begin = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.march_ms()  # drop the warm-up launches
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    launches = scft_b200.launch_count()
    ms = ev0.elapsed_time(ev1)
    march_tot, march_cnt = eng.march_ms()
    eng.set_timing(False)
    t = torch.tensor([ms, march_tot / max(march_cnt, 1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, march_ms_avg = float(t[0]), float(t[1])

    # ---- end to end through the C ABI with host buffers (H2D + kernel + D2H inside the call)
    barrier()
    for _ in range(2):
        eng.residual_host_ptr(P, h_eta.data_ptr(), h_out.data_ptr())
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.residual_host_ptr(P, h_eta.data_ptr(), h_out.data_ptr())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te[0])

    # per-problem scalars of the whole sweep: the one collective of the workload, outside the timed region
    from scft_b200 import sweep
    done, iters, err = mixer.status(stream.cuda_stream)
    local = np.stack([err, iters.astype(np.float64)], axis=1)
    allres = sweep.gather_results(local, world * P, rank, world) if world > 1 else local
    if rank == 0:
        assert allres.shape[0] == world * P
        finite = int(np.isfinite(allres[:, 0]).sum())
        dof_steps_per_step = world * P * ni * NSTEPS
        value = dof_steps_per_step * args.steps / (ms_total * 1e-3)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        algo_bytes = P * ni * NSTEPS * BYTES_PER_DOF_STEP
        traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum of one march launch (ncu --set full capture)
        tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if os.path.exists(tpath):
            tr = json.load(open(tpath))["march_ie_kernel"]
            if tr["problems"] == P and tr["N"] == N_NODES and tr["nsteps"] == NSTEPS:
                traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
        achieved = algo_bytes / (march_ms_avg * 1e-3) / 1e9
        line = {"metric": "propagator_dof_steps_per_s", "value": value, "unit": "DOF-steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(P),
                "scft_iterations_per_s": world * P * args.steps / (ms_total * 1e-3),
                "problems_with_finite_residual_at_end": finite,
                "clocks": clocks,
                "e2e": {"value": dof_steps_per_step * args.steps / e2e_s, "unit": "DOF-steps/s",
                        "h2d_bytes_per_step": P * ni * 8, "d2h_bytes_per_step": P * ni * 8,
                        "call": "scftb_residual_batch (pinned host buffers), wall clock, max over ranks"},
                "gpu_launches": launches,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "kernel": "march_ie_kernel<8,128,uniform>",
                             "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": march_ms_avg,
                             "peak_source": peak_src,
                             "bytes_per_dof_step": BYTES_PER_DOF_STEP}}
        if not args.no_cpu_baseline and world == 1:
            cores = host_cores()
            cnt = max(cores, 8) * 256   # ~10-20 s of CPU work (16 threads: 4096 problems in ~11 s)
            v, dt = oracle_throughput(cnt, cores)
            line["cpu_baseline"] = {"value": v, "unit": "DOF-steps/s", "cores": cores, "kind": "port",
                                    "sample": f"first {cnt} problems of the sweep, oracle/scft_oracle.c, {dt:.1f} s"}
            probe = reference_code_probe()
            if probe:
                line["cpu_baseline"]["reference_code_probe"] = probe
        _emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
