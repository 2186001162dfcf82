#!/usr/bin/env python
"""bench.py -- the headline metric of BASELINE.json on B200.

Workload (N = 1): BASELINE config 2, ``Toeplitz{Float64}`` n = 2^20 times a matrix of 1024
right-hand-side columns (``mul!(C, factorize(A), B)``, /root/reference/src/linearalgebra.jl:88-99).
A "step" is one pass of the hot path over the whole batch.  With N GPUs every rank owns its
own 1024 columns (weak scaling, no data-path collective; the spectrum is broadcast once per
factorization over NCCL before the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  ``value`` = matvecs/s with inputs resident in HBM, timed on
the device with CUDA events; ``e2e`` = the same metric through ``tb200_mul_host`` with pinned
HOST buffers (H2D + compute + D2H inside the timed region); ``roofline`` = algorithmic bytes
of the dominant kernel / its CUDA-event duration vs the measured HBM peak; ``cpu_baseline`` =
the CPU restatement of the reference algorithm (oracle/, FFT length m+n-1, pocketfft) on a
bounded column sample using every host core.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LOG2N = 20
NCOLS = 1024
ALG_BYTES_PER_MATVEC = 2 * (1 << LOG2N) * 8          # read x (n*8) + write y (n*8) = 16 MiB
METRIC = "batched Toeplitz matvec matvecs/s (n=2^20, Float64)"
UNIT = "matvecs/s"


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._pump, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------
# CPU baseline: the oracle's restatement of the reference path (bounded sample)
# ---------------------------------------------------------------------------------------
def cpu_reference_columns(vc, vr, Xs, cores):
    """mul!(C, factorize(A), B) on a column sample, reference algorithm (N = m+n-1 FFTs),
    columns spread over `cores` threads (each column is the reference's serial body)."""
    import numpy as np
    from oracle import toeplitz_oracle as O
    A = O.Toeplitz(vc, vr)
    t0 = time.perf_counter()
    F = O.factorize(A)
    Y = np.zeros_like(Xs, order="F")

    def one(j):
        O.mul_factorization(Y[:, j], F, Xs[:, j], True, False)

    with cf.ThreadPoolExecutor(max_workers=cores) as ex:
        list(ex.map(one, range(Xs.shape[1])))
    return Y, time.perf_counter() - t0


def coefficient_vectors():
    import numpy as np
    n = 1 << LOG2N
    k = np.arange(n, dtype=np.float64)
    return 0.9 ** k, 0.4 ** k          # the reference's test recipe (test/runtests.jl:24-25)


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the Julia reference cannot run
    here) on this box's host cores; each step = one bounded column sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    cores = host_cores()
    sample = max(cores, 16)
    vc, vr = coefficient_vectors()
    rng = np.random.default_rng(0)
    Xs = np.asfortranarray(rng.standard_normal((1 << LOG2N, sample)))
    for _ in range(max(args.warmup, 0) and 1):
        cpu_reference_columns(vc, vr, Xs, cores)
    t = 0.0
    for _ in range(args.steps):
        _, dt = cpu_reference_columns(vc, vr, Xs, cores)
        t += dt
    val = sample * args.steps / t
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "Toeplitz{Float64} n=2^20 x matrix, %d-column sample per step (of 1024)" % sample,
                   "algorithm": "reference path restated (oracle/): factorize + per-column fft/ifft of length m+n-1, pocketfft"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d columns per step, %d steps, columns spread over %d threads" % (sample, args.steps, cores)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import toeplitz_b200 as tb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << LOG2N
    ncols = args.cols
    vc, vr = coefficient_vectors()
    if args.group_mb:
        tb.set_option("group_bytes", args.group_mb << 20)
    if args.no_overlap:
        tb.set_option("overlap_streams", 1)
    for kv in args.opt:
        k, v = kv.split("=")
        tb.set_option(k, int(v))

    # factorize on rank 0, broadcast the spectrum once (the only exchange on the path)
    if rank == 0:
        F = tb.factorize(tb.Toeplitz(torch.from_numpy(vc).cuda(), torch.from_numpy(vr).cuda()))
    else:
        F = tb.factorize_empty_like(0, torch.float64, (n, n))
    if world > 1:
        torch.cuda.synchronize()
        dist.broadcast(tb.spectrum_tensor(F), src=0)
    torch.cuda.synchronize()

    g = torch.Generator(device="cuda")
    g.manual_seed(1234 + rank)
    X = torch.randn((ncols, n), dtype=torch.float64, device="cuda", generator=g).t()     # column-major n x ncols
    Y = tb.colmajor(n, ncols, torch.float64)

    def step():
        tb.mul_(Y, F, X, True, False)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    tb.set_option("time_passes", 0)
    tb.pass_times()                     # drain
    l0 = tb.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    launches = tb.launch_count() - l0
    # ---- second timed region of K steps for the roofline leg: every pass bracketed by CUDA events on its
    # launching stream, group overlap off so a pass duration is one kernel's duration (the events and the lost
    # overlap cost ~40 %, which is why `value` comes from the first region)
    pass_ms = {"cols_A": 0.0, "rows_B": 0.0, "cols_C": 0.0, "rows_single": 0.0}
    pass_cnt = {k: 0 for k in pass_ms}
    if not args.no_pass_events:
        tb.set_option("time_passes", 1)
        tb.set_option("overlap_streams", 1)
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
        pass_ms, pass_cnt = tb.pass_times()
        tb.set_option("time_passes", 0)
        tb.set_option("overlap_streams", 2)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * ncols * args.steps / (ms * 1e-3)

    # ---- end-to-end through the host-buffer C-ABI call (pinned host in, pinned host out)
    e2e_steps = max(1, min(args.steps, 3))
    # host staging is pinned memory: bound it per rank when several ranks share one host (8 x 16 GiB would not be polite)
    e2e_cols = ncols if world == 1 else min(ncols, 512)
    if args.no_e2e:
        if world > 1:
            dist.destroy_process_group()
        if rank == 0:
            print(json.dumps({"value": value, "ms_per_step": ms / args.steps, "pass_ms": pass_ms, "pass_cnt": pass_cnt,
                              "launches": launches, "clocks": clocks}), flush=True)
        return
    Xh = torch.empty((e2e_cols, n), dtype=torch.float64).pin_memory().t()
    Yh = torch.empty((e2e_cols, n), dtype=torch.float64).pin_memory().t()
    Xh.copy_(X[:, :e2e_cols])
    torch.cuda.synchronize()
    tb.mul_host_(Yh, F, Xh)             # warm-up (allocates staging buffers, touches pages)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        tb.mul_host_(Yh, F, Xh)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = world * e2e_cols * e2e_steps / e2e_s
    e2e_err = float((Yh[:, :2] - Y[:, :2].cpu()).abs().max())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (CUDA events recorded around every pass, live)
    peak, peak_src = measured_hbm_peak()
    dom = max(pass_ms, key=lambda k: pass_ms[k])
    per_launch_ms = pass_ms[dom] / max(pass_cnt[dom], 1)
    cols_per_launch = ncols * args.steps / max(pass_cnt[dom], 1)
    achieved = (ALG_BYTES_PER_MATVEC * cols_per_launch / (per_launch_ms * 1e-3) / 1e9) if per_launch_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(dom)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "alg_bytes_per_launch": ALG_BYTES_PER_MATVEC * cols_per_launch, "avg_launch_ms": per_launch_ms,
                "pass_ms_total": pass_ms, "pass_launches": pass_cnt,
                "timed_how": "CUDA events around every pass on its launching stream, second K-step region, overlap off",
                "whole_step_frac": ALG_BYTES_PER_MATVEC * ncols / (ms / args.steps * 1e-3) / 1e9 / peak,
                "co_bound": "FP64 FMA pipe: ~210 MFLOP per matvec (SURVEY 8d)"}
    # the same step seen from the FP64 pipe: nominal 5 N log2 N flops per complex transform of N = 2^21, two transforms
    # per column pair; peak = FMA throughput measured on this pool (profiles/tools/fp64_peak.cu), of which an FFT's
    # add-dominated butterflies can use about half
    fp64_flop_per_matvec = 5.0 * (2 * n) * math.log2(2 * n)
    fp64_tflops = fp64_flop_per_matvec * ncols / (ms / args.steps * 1e-3) / 1e12
    roofline["fp64"] = {"flop_per_matvec": fp64_flop_per_matvec, "achieved_tflops": fp64_tflops, "peak_tflops": 33.7,
                        "frac": fp64_tflops / 33.7}

    # ---- CPU baseline on a bounded sample of the same workload, rank 0 only
    cpu = None
    parity = None
    if not args.no_cpu:
        cores = host_cores()
        sample = max(cores, 16) * 4            # ~17 CPU-seconds of work
        Xs = np.asfortranarray(X[:, :sample].cpu().numpy())
        Yo, dt = cpu_reference_columns(vc, vr, Xs, cores)
        cpu = {"value": sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "first %d of the %d columns, reference algorithm (FFT length 2^21-1, pocketfft), "
                         "columns spread over %d threads, %.1f s" % (sample, ncols, cores, dt)}
        Yg = Y[:, :sample].cpu().numpy()
        parity = float(np.linalg.norm(Yg - Yo) / np.linalg.norm(Yo))

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "Toeplitz{Float64} n=2^20 x Matrix with %d RHS columns per GPU (BASELINE configs[1])" % ncols,
                   "n": n, "columns_per_gpu": ncols, "parallelism": "columns sharded over %d GPU(s), spectrum broadcast once" % world,
                   "l2": "inputs (8 GiB X + 8 GiB Y per GPU) far exceed the 126 MB L2; no flush needed"},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": e2e_cols * n * 8, "d2h_bytes_per_step": e2e_cols * n * 8,
                "steps": e2e_steps, "columns_per_gpu_per_step": e2e_cols, "max_abs_diff_vs_device_path": e2e_err},
        "gpu_launches": int(launches),
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "parity_rel_err_vs_oracle": parity,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cols", type=int, default=NCOLS, help="columns per GPU (default: the BASELINE config, 1024)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-pass-events", action="store_true", help="do not bracket each pass with CUDA events")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg")
    ap.add_argument("--group-mb", type=int, default=0, help="tuning: two-level workspace budget per group (MiB)")
    ap.add_argument("--no-overlap", action="store_true", help="tuning: disable the two-stream group overlap")
    ap.add_argument("--opt", action="append", default=[], help="tuning: library option name=value (repeatable)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
