#!/usr/bin/env python
"""bench.py -- the headline metric of BASELINE.json on B200, plus every other BASELINE config.

Headline workload (the line's `value`): BASELINE config 2, ``Toeplitz{Float64}`` n = 2^20 times a matrix of 1024
right-hand-side columns (``mul!(C, factorize(A), B)``, /root/reference/src/linearalgebra.jl:88-99).  A "step" is one
pass of the hot path over the whole batch.  With N GPUs:

* ``value``            weak scaling: every rank owns its own 1024 columns (no data-path collective; the spectrum is
                       broadcast once per factorization over NCCL before the timed region);
* ``strong_scaling``   the BASELINE wording itself: 1024 columns IN TOTAL, sharded with ``sharding.column_shard``;
* ``configs``          C1 (single vector, us), C3 (Circulant ldiv! solves/s), C4 (Levinson / Durbin solves/s),
                       C5a (Hankel matvec), C5b (cgs + Strang iterations/s), each with its own roofline fraction on
                       SURVEY 8(d)'s algorithmic bytes and a parity check; C3 and C4 shard over the ranks (strong),
                       the single-vector configs run on rank 0 ("replicas only");
* ``parity``           relative error vs the CPU oracle on columns of EVERY rank (after the broadcast), max-reduced.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  ``value`` = matvecs/s with inputs resident in HBM, timed on the device with CUDA
events; ``e2e`` = the same metric through ``tb200_mul_host`` with pinned HOST buffers (H2D + compute + D2H inside the
timed region; same column count at every N; a pageable-host figure beside it); ``roofline`` = algorithmic bytes of
the dominant kernel / its CUDA-event duration vs the measured HBM peak; ``cpu_baseline`` = the CPU restatement of the
reference algorithm (oracle/, FFT length m+n-1, pocketfft) on a bounded column sample using every host core.
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
E2E_COLS = 512                                      # columns per host-buffer call (8 GiB of pinned host per rank); a step makes ncols / 512 calls
ALG_BYTES_PER_MATVEC = 2 * (1 << LOG2N) * 8          # read x (n*8) + write y (n*8) = 16 MiB
METRIC = "batched Toeplitz matvec matvecs/s (n=2^20, Float64)"
UNIT = "matvecs/s"
FP64_PEAK_TFLOPS = 33.7                              # measured on this pool with profiles/tools/fp64_peak.cu (round 1)
FP32_PEAK_TFLOPS = 70.4


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


def bind_to_gpu_numa(local: int):
    """Pin this rank to the cores of its GPU's NUMA node before any host buffer is allocated, so pinned staging
    memory is first-touched next to the PCIe root complex it will stream through.  Best effort."""
    info = {"numa_node": None, "cores": None}
    try:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = vis.split(",")[local] if vis else str(local)
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", idx],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if not bus:
            return info
        dom, rest = bus.split(":", 1)
        path = "/sys/bus/pci/devices/%s:%s/numa_node" % (dom[-4:], rest)
        node = int(open(path).read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        cl = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
        cpus = set()
        for part in cl.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cores"] = len(allowed)
    except Exception as e:
        info["error"] = str(e)[:80]
    return info


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
# helpers of our arm
# ---------------------------------------------------------------------------------------
def timed_region(fn, reps, dist=None, world=1):
    """K calls between a barrier + synchronize on both sides, CUDA events, max over ranks -> seconds per call."""
    import torch
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms * 1e-3 / reps


def rel_err(a, b):
    import numpy as np
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def max_over_ranks(v, dist, world):
    import torch
    if world == 1:
        return float(v)
    t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_configs(tb, rank, world, dist, peak, quick=False):
    """The other BASELINE configs: device-resident inputs, CUDA-event timing, parity sample per config."""
    import numpy as np
    import torch
    from toeplitz_b200 import sharding
    from oracle import toeplitz_oracle as O
    from oracle import build_oracle as OC
    out = {}
    eps = np.finfo(np.float64).eps

    # ---- C1: Toeplitz{Float64} n = 2^16, one vector (rank 0; latency-bound, "replicas only") ----------------------------
    if rank == 0:
        n = 1 << 16
        k = np.arange(n, dtype=np.float64)
        vc, vr = 0.9 ** k, 0.4 ** k
        A = tb.Toeplitz(torch.from_numpy(vc).cuda(), torch.from_numpy(vr).cuda())
        xh = np.random.default_rng(0).standard_normal(n)
        x = torch.from_numpy(xh).cuda()
        y = torch.empty_like(x)
        F = tb.factorize(A)
        for _ in range(3):
            tb.mul_(y, F, x); tb.mul_(y, A, x)
        t_mul = timed_region(lambda: tb.mul_(y, F, x), 200)
        # the same call without the Python mirror's argument checks (what a `ccall` from the Julia extension pays), and as
        # a captured CUDA graph (device time of the three dependent launches alone)
        from toeplitz_b200 import _lib as L
        from toeplitz_b200.api import _DT
        lib, one, zero, dt = L.lib(), L.dbl2(1.0), L.dbl2(0.0), _DT[torch.float64]
        xp, yp, st0 = x.data_ptr(), y.data_ptr(), torch.cuda.current_stream().cuda_stream
        t_cabi = timed_region(lambda: lib.tb200_mul(F._h, yp, n, dt, xp, n, dt, 1, one, zero, st0), 200)
        t_graph = None
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                tb.mul_(y, F, x)
                side.synchronize()
                with torch.cuda.graph(graph, stream=side):
                    tb.mul_(y, F, x)
            graph.replay()
            t_graph = timed_region(graph.replay, 200)
        except Exception as e:                                   # a number, not a requirement of the bench
            print("C1 graph replay skipped: %s" % str(e)[:120], file=sys.stderr)
        t_fac = timed_region(lambda: tb.factorize(A), 50)
        t_ref = timed_region(lambda: tb.mul_(y, A, x), 50)             # reference semantics: re-factorize on every call (:37)
        par = rel_err(y.cpu().numpy(), O.matvec(O.Toeplitz(vc, vr), xh))
        out["C1"] = {"workload": "Toeplitz{Float64} n=2^16, single vector (BASELINE configs[0])", "mul_us": t_mul * 1e6,
                     "mul_us_c_abi": t_cabi * 1e6, "mul_us_graph_replay": None if t_graph is None else t_graph * 1e6,
                     "note": "mul_us goes through the Python mirror of the reference API (about 1 us of host time more than the "
                             "device needs per call); mul_us_c_abi is tb200_mul called directly, mul_us_graph_replay the device "
                             "time of the three dependent launches",
                     "factorize_us": t_fac * 1e6, "mul_with_refactorize_us": t_ref * 1e6,
                     "roofline": {"bound": "latency (3 launches)", "achieved": 2 * n * 8 / t_mul / 1e9, "peak": peak, "unit": "GB/s",
                                  "frac": 2 * n * 8 / t_mul / 1e9 / peak, "alg_bytes_per_unit": 2 * n * 8},
                     "parity_rel_err_vs_oracle": par, "tolerance": 1e-12}
        del F, A, x, y

    # ---- C3: Circulant{ComplexF64} n = 2^22, 256 RHS ldiv!, sharded by columns -----------------------------------------
    n, l = 1 << 22, (64 if quick else 256)
    a, b = sharding.column_shard(l, world, rank)
    lc = b - a
    vch = (0.9 ** np.arange(n)).astype(np.complex128)
    C = tb.Circulant(torch.from_numpy(vch).cuda())
    t_fac = timed_region(lambda: tb.factorize(C), 10) if rank == 0 else 0.0
    F = tb.factorize(C)
    g = torch.Generator(device="cuda")
    g.manual_seed(200 + rank)
    B = torch.randn((lc, n), dtype=torch.complex128, device="cuda", generator=g).t()
    b0 = B[:, 0].clone().cpu().numpy()
    blast = B[:, lc - 1].clone().cpu().numpy()
    X = tb.colmajor(n, lc, torch.complex128)
    tb.circ_ldiv_(F, B, out=X)                                      # warm-up (workspaces, reciprocal spectrum)
    t_ld = timed_region(lambda: tb.circ_ldiv_(F, B, out=X), 3, dist, world)
    par = max(rel_err(X[:, 0].cpu().numpy(), O.circ_ldiv(O.Circulant(vch), b0.copy())),
              rel_err(X[:, lc - 1].cpu().numpy(), O.circ_ldiv(O.Circulant(vch), blast.copy())))
    par = max_over_ranks(par, dist, world)
    t_eig = timed_region(lambda: tb.eigvals(F), 20) if rank == 0 else 0.0
    # C*(C\\B) == B on the whole shard (size-independent property)
    R = tb.colmajor(n, lc, torch.complex128)
    tb.mul_(R, F, X)
    prop = max_over_ranks(float(torch.linalg.norm(R - B) / torch.linalg.norm(B)), dist, world)
    if rank == 0:
        gbs = l * 2 * n * 16 / t_ld / 1e9
        out["C3"] = {"workload": "Circulant{ComplexF64} n=2^22, %d RHS ldiv! sharded over %d GPU(s) (BASELINE configs[2])" % (l, world),
                     "ldiv_solves_per_s": l / t_ld, "ms_per_batch": t_ld * 1e3, "factorize_ms": t_fac * 1e3, "eigvals_ms": t_eig * 1e3,
                     "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak * world, "unit": "GB/s", "frac": gbs / (peak * world),
                                  "alg_bytes_per_unit": 2 * n * 16, "co_bound": "FP64 FMA: ~923 MFLOP per solve"},
                     "parity_rel_err_vs_oracle": par, "roundtrip_C_times_CinvB_rel_err": prop, "tolerance": 1e-12}
    del B, X, R, F, C
    torch.cuda.empty_cache()

    # ---- C4: SymmetricToeplitz{Float64} n = 512, 10^6 systems, Levinson + Durbin, sharded by systems ------------------
    n, nsys = 512, (100000 if quick else 1000000)
    a, b = sharding.column_shard(nsys, world, rank)
    ns = b - a
    g.manual_seed(300 + rank)
    Rm = (torch.rand((ns, n), dtype=torch.float64, device="cuda", generator=g) / n).t()        # r = u/n (test/runtests.jl:117)
    Bm = torch.randn((ns, n), dtype=torch.float64, device="cuda", generator=g).t()
    Xm = tb.colmajor(n, ns)
    Rl = Rm[: n - 1, :]
    tb.levinson_(Xm, Rl, Bm)                                        # warm-up
    t_lev = timed_region(lambda: tb.levinson_(Xm, Rl, Bm), 3, dist, world)
    samp = list(range(0, ns, max(1, ns // 32)))[:32]
    Rs = np.asfortranarray(Rl[:, samp].cpu().numpy())
    Bs = np.asfortranarray(Bm[:, samp].cpu().numpy())
    par_l = rel_err(Xm[:, samp].cpu().numpy(), OC.c_levinson_batched(Rs, Bs))
    Ym = tb.colmajor(n, ns)
    tb.durbin_(Ym, Rm)                                              # warm-up
    t_dur = timed_region(lambda: tb.durbin_(Ym, Rm), 3, dist, world)
    Rd = np.asfortranarray(Rm[:, samp].cpu().numpy())
    par_d = rel_err(Ym[:, samp].cpu().numpy(), OC.c_durbin_batched(Rd))
    par_l, par_d = max_over_ranks(par_l, dist, world), max_over_ranks(par_d, dist, world)
    if rank == 0:
        # algorithmic flops of the reference recursion (5k FMAs per step: 2.5 n^2 FMAs = ~4 n^2 flop with its divisions);
        # the generator-form kernels execute 3 n^2 (Levinson) / 2 n^2 (Durbin) FMAs per solve -- both are reported
        tf_l, tf_d = 4.0 * n * n * nsys / t_lev / 1e12, 2.0 * n * n * nsys / t_dur / 1e12
        ex_l, ex_d = 6.0 * n * n * nsys / t_lev / 1e12, 4.0 * n * n * nsys / t_dur / 1e12
        out["C4"] = {"workload": "SymmetricToeplitz{Float64} n=512, %d independent systems sharded over %d GPU(s) (BASELINE configs[3])" % (nsys, world),
                     "levinson_solves_per_s": nsys / t_lev, "durbin_solves_per_s": nsys / t_dur,
                     "roofline": {"bound": "fp64", "achieved": tf_l, "peak": FP64_PEAK_TFLOPS * world, "unit": "TFLOP/s",
                                  "frac": tf_l / (FP64_PEAK_TFLOPS * world), "durbin_achieved": tf_d,
                                  "durbin_frac": tf_d / (FP64_PEAK_TFLOPS * world),
                                  "executed_fma_tflops": {"levinson": ex_l, "durbin": ex_d},
                                  "executed_frac": {"levinson": ex_l / (FP64_PEAK_TFLOPS * world), "durbin": ex_d / (FP64_PEAK_TFLOPS * world)},
                                  "peak_source": "FP64 FMA rate measured with profiles/tools/fp64_peak.cu (round 1), per GPU",
                                  "alg_bytes_per_unit": 12280, "alg_gbs": nsys * 12280 / t_lev / 1e9},
                     "parity_rel_err_vs_oracle": {"levinson": par_l, "durbin": par_d}, "tolerance": 1e-12}
    del Rm, Bm, Xm, Ym, Rl
    torch.cuda.empty_cache()

    # ---- C5a / C5b: single vectors at n = 2^24, ComplexF32 (rank 0; "replicas only") ------------------------------------
    if rank == 0:
        n = 1 << (22 if quick else 24)
        g.manual_seed(500)
        v = torch.randn(2 * n - 1, dtype=torch.complex64, device="cuda", generator=g)
        x = torch.randn(n, dtype=torch.complex64, device="cuda", generator=g)
        y = torch.empty_like(x)
        F = tb.factorize(tb.Hankel(v, (n, n)))
        for _ in range(3):
            tb.mul_(y, F, x)
        t_h = timed_region(lambda: tb.mul_(y, F, x), 20)
        # definition-level check on sampled rows: y[i] = sum_j v[i+j] x[j] in complex128
        vh, xh, yh = v.cpu().numpy().astype(np.complex128), x.cpu().numpy().astype(np.complex128), y.cpu().numpy()
        rows = [0, 1, n // 3, n // 2 + 17, n - 2, n - 1]
        ref_rows = np.array([np.dot(vh[i:i + n], xh) for i in rows])
        par_h = rel_err(yh[rows], ref_rows)
        gbs = 2 * n * 8 / t_h / 1e9
        out["C5a"] = {"workload": "Hankel{ComplexF32} n=2^%d matvec via the reversed-index embedding (BASELINE configs[4])" % int(math.log2(n)),
                      "matvecs_per_s": 1.0 / t_h, "ms": t_h * 1e3,
                      "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                   "alg_bytes_per_unit": 2 * n * 8, "co_bound": "FP32 FMA: ~8.4 GFLOP per matvec"},
                      "parity_rel_err_vs_dense_rows": par_h, "tolerance": 1e-5}
        del F, v, x, y
        torch.cuda.empty_cache()
        vck = (0.5 ** np.arange(n)).astype(np.complex64)
        A = tb.SymmetricToeplitz(torch.from_numpy(vck).cuda())
        FA = tb.factorize(A)
        Mp = tb.factorize(tb.strang(A))
        bvec = torch.randn(n, dtype=torch.complex64, device="cuda", generator=g)
        its = 20
        tb.cgs(FA, torch.zeros_like(bvec), bvec, Mp, 2, 0.0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, err, it, flag = tb.cgs(FA, torch.zeros_like(bvec), bvec, Mp, its, 0.0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        # run to the reference's tolerance and check the TRUE residual on sampled rows of the dense definition
        xs, err_c, it_c, flag_c = tb.cgs(FA, torch.zeros_like(bvec), bvec, Mp, 1000, 1e-5)
        xs_h, b_h = xs.cpu().numpy().astype(np.complex128), bvec.cpu().numpy().astype(np.complex128)
        col = (0.5 ** np.arange(64)).astype(np.float64)               # 0.5^k < 1e-19 beyond k = 64: banded to working precision
        res_rows = []
        for i in (0, 5, n // 2, n - 6, n - 1):
            lo, hi = max(0, i - 63), min(n, i + 64)
            res_rows.append(b_h[i] - np.dot(col[np.abs(np.arange(lo, hi) - i)], xs_h[lo:hi]))
        res_s = float(np.max(np.abs(res_rows)) / (np.linalg.norm(b_h) / math.sqrt(n)))
        # the norm the solver reports, recomputed from scratch on the device: ||b - A x|| / ||b||
        r_dev = bvec - tb.mul(FA, xs)
        res_true = float(torch.linalg.vector_norm(r_dev) / torch.linalg.vector_norm(bvec))
        gbs = 19 * n * 8 * it / dt / 1e9
        out["C5b"] = {"workload": "SymmetricToeplitz{ComplexF32} n=2^%d, cgs + Strang preconditioner, %d timed iterations (BASELINE configs[4])" % (int(math.log2(n)), it),
                      "iterations_per_s": it / dt, "ms_per_iteration": dt / it * 1e3,
                      "converged": {"iterations": it_c, "flag": flag_c, "rel_residual_reported": err_c, "tol": 1e-5,
                                    "true_rel_residual_recomputed": res_true,
                                    "max_abs_residual_on_sampled_rows_over_rms_b": res_s,
                                    "note": "rows 0, 5, n/2, n-6, n-1 of the dense definition in complex128; after two Strang-preconditioned "
                                            "iterations the residual sits in the few boundary rows, so its largest entries are O(1e-2) "
                                            "while its 2-norm over 2^24 rows is below tol"},
                      "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                   "alg_bytes_per_unit": 19 * n * 8, "note": "19 n-vector sweeps per iteration (SURVEY 8d) + 8 FFT applications"}}
        del FA, Mp, A, bvec, xs
        torch.cuda.empty_cache()
    return out


def pcie_probe(gib=1.0):
    """Raw pinned-copy ceiling of this rank's host link: H2D and D2H at once (what the e2e leg is bounded by)."""
    import torch
    nbytes = int(gib * (1 << 30))
    h1 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d1 = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d2 = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(2):
        with torch.cuda.stream(s1):
            d1.copy_(h1, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        with torch.cuda.stream(s1):
            d1.copy_(h1, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return 2 * 3 * nbytes / dt / 1e9


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    all_cores = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa(local)
    import toeplitz_b200 as tb
    from toeplitz_b200 import sharding
    from oracle import toeplitz_oracle as O

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

    # factorize on rank 0, broadcast the spectrum once (the only exchange on the path) through the C ABI's own NCCL
    # communicator (tb200_comm_init_rank / tb200_bcast_spectrum)
    if rank == 0:
        F = tb.factorize(tb.Toeplitz(torch.from_numpy(vc).cuda(), torch.from_numpy(vr).cuda()))
    else:
        F = tb.factorize_empty_like(0, torch.float64, (n, n))
    bcast_how = "none (1 GPU)"
    if world > 1:
        torch.cuda.synchronize()
        try:
            comm = sharding.Communicator.from_torch_group()
            comm.bcast_spectrum(F, root=0)
            torch.cuda.synchronize()
            comm.destroy()
            bcast_how = "tb200_bcast_spectrum (C ABI, ncclBroadcast)"
        except Exception as e:                                   # keep the bench alive: torch's NCCL does the same broadcast
            dist.broadcast(tb.spectrum_tensor(F), src=0)
            tb.spectrum_ready(F)
            bcast_how = "torch.distributed.broadcast (C-ABI communicator failed: %s)" % str(e)[:80]
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
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    tb.set_option("time_passes", 0)
    tb.pass_times()                     # drain
    l0 = tb.launch_count()
    sec = timed_region(step, args.steps, dist, world)
    ms = sec * 1e3 * args.steps
    launches = tb.launch_count() - l0
    value = world * ncols * args.steps / (ms * 1e-3)

    # ---- parity of THIS rank's output vs the CPU oracle (non-root ranks run on the broadcast spectrum) ------------------
    pcols = [0, ncols - 1] if ncols > 1 else [0]
    Xp = np.asfortranarray(X[:, pcols].cpu().numpy())
    Yo, _ = cpu_reference_columns(vc, vr, Xp, min(host_cores(), len(pcols)))
    parity_rank = rel_err(Y[:, pcols].cpu().numpy(), Yo)
    # checksum of the whole batch against a second product of the same inputs through the single-stream path
    parity_all = max_over_ranks(parity_rank, dist, world)
    # every column of the full-size batch at once, by linearity: (A X) w == A (X w) for a seeded random weight vector w
    # (distinct weights, so a column permutation or a dropped column shows up as an O(1) error); the right-hand side is
    # one more product through the library, so the check costs one matvec plus two passes over X and Y
    w = torch.randn(ncols, dtype=torch.float64, device="cuda", generator=g)
    yw = tb.mul(F, torch.mv(X, w))
    Yw = torch.mv(Y, w)
    lin_rank = float(torch.linalg.vector_norm(Yw - yw) / torch.linalg.vector_norm(yw))
    lin_all = max_over_ranks(lin_rank, dist, world)
    del w, yw, Yw

    # ---- strong scaling: 1024 columns IN TOTAL, column-sharded (the BASELINE wording) -----------------------------------
    strong = None
    if world > 1 and not args.no_strong:
        a, b = sharding.column_shard(NCOLS, world, rank)
        Xs_, Ys_ = X[:, : b - a], Y[:, : b - a]
        for _ in range(3):
            tb.mul_(Ys_, F, Xs_)
        sec_s = timed_region(lambda: tb.mul_(Ys_, F, Xs_), args.steps, dist, world)
        strong = {"value": NCOLS / sec_s, "unit": UNIT, "ms_per_step": sec_s * 1e3, "columns_total": NCOLS,
                  "columns_per_gpu": b - a, "scaling": "strong",
                  "whole_step_frac": ALG_BYTES_PER_MATVEC * NCOLS / sec_s / 1e9 / (measured_hbm_peak()[0] * world)}

    # ---- second timed region of K steps for the roofline leg: every pass bracketed by CUDA events on its
    # launching stream, group overlap off so a pass duration is one kernel's duration (the events and the lost
    # overlap cost ~40 %, which is why `value` comes from the first region)
    pass_ms = {"cols_A": 0.0, "rows_B": 0.0, "cols_C": 0.0, "rows_single": 0.0, "cols_CA": 0.0}
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

    if args.no_e2e:
        if world > 1:
            dist.destroy_process_group()
        if rank == 0:
            print(json.dumps({"value": value, "ms_per_step": ms / args.steps, "pass_ms": pass_ms, "pass_cnt": pass_cnt,
                              "launches": launches, "clocks": clocks, "parity": parity_all, "strong": strong}), flush=True)
        return

    # ---- end-to-end through the host-buffer C-ABI call (pinned host in, pinned host out), E2E_COLS per GPU at every N -----
    e2e_steps = max(1, min(args.steps, 3))
    e2e_cols = min(ncols, E2E_COLS)
    Xh = torch.empty((e2e_cols, n), dtype=torch.float64).pin_memory().t()
    Yh = torch.empty((e2e_cols, n), dtype=torch.float64).pin_memory().t()
    Xh.copy_(X[:, :e2e_cols])
    torch.cuda.synchronize()
    tb.mul_host_(Yh, F, Xh)             # warm-up (allocates staging buffers, touches pages)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_calls = max(1, -(-ncols // e2e_cols))       # a step = all ncols columns of the config, through the same pinned matrices
    for _ in range(e2e_steps):
        for _c in range(e2e_calls):
            tb.mul_host_(Yh, F, Xh)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0, dist, world)
    e2e_val = world * e2e_cols * e2e_calls * e2e_steps / e2e_s
    e2e_err = float((Yh[:, :2] - Y[:, :2].cpu()).abs().max())
    link = pcie_probe(1.0)
    if world > 1:
        t = torch.tensor([link], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        link_sum = float(t.item())
    else:
        link_sum = link
    # pageable host matrices (what a plain Julia Matrix is): a quarter of the columns, rank 0 only
    pageable = None
    if rank == 0:
        pc = max(2, e2e_cols // 4)
        Xp_ = torch.empty((pc, n), dtype=torch.float64).t()
        Yp_ = torch.empty((pc, n), dtype=torch.float64).t()
        Xp_.copy_(X[:, :pc])
        tb.mul_host_(Yp_, F, Xp_)
        t0 = time.perf_counter()
        tb.mul_host_(Yp_, F, Xp_)
        torch.cuda.synchronize()
        dtp = time.perf_counter() - t0
        pageable = {"value": pc / dtp, "unit": UNIT, "columns": pc, "host_gbs": 2 * pc * n * 8 / dtp / 1e9,
                    "note": "pageable (unpinned) host matrices on rank 0 while the other ranks idle"}
        del Xp_, Yp_
    del Xh, Yh
    try:
        os.sched_setaffinity(0, all_cores)          # the CPU legs below use every host core again
    except Exception:
        pass

    # ---- the other BASELINE configs ------------------------------------------------------------------------------------
    peak, peak_src = measured_hbm_peak()
    configs = None
    if not args.no_configs:
        del X, Y
        torch.cuda.empty_cache()
        configs = run_configs(tb, rank, world, dist, peak, quick=args.quick_configs)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (CUDA events recorded around every pass, live)
    dom = max(pass_ms, key=lambda k: pass_ms[k])
    per_launch_ms = pass_ms[dom] / max(pass_cnt[dom], 1)
    cols_per_launch = ncols * args.steps / max(pass_cnt[dom], 1)
    achieved = (ALG_BYTES_PER_MATVEC * cols_per_launch / (per_launch_ms * 1e-3) / 1e9) if per_launch_ms > 0 else 0.0
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get(dom)
            traffic_src = "static: ncu capture committed under profiles/ (%s), not measured in this run" % tj.get("_source", "traffic.json")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "alg_bytes_per_launch": ALG_BYTES_PER_MATVEC * cols_per_launch, "avg_launch_ms": per_launch_ms,
                "pass_ms_total": pass_ms, "pass_launches": pass_cnt,
                "timed_how": "CUDA events around every pass on its launching stream, second K-step region, overlap off",
                "whole_step_frac": ALG_BYTES_PER_MATVEC * ncols / (ms / args.steps * 1e-3) / 1e9 / peak,
                "co_bound": "FP64 FMA pipe: ~210 MFLOP per matvec (SURVEY 8d)"}
    # the second fraction -- DRAM bytes the kernel really moves (committed ncu counters) per second of its
    # launch, against the same peak; "cold" = every launch starts with a flushed L2 (ncu default), "warm" = as in the loop
    try:
        tjw = json.load(open(tpath)).get("warm", {})
        if traffic and per_launch_ms > 0:
            roofline["dram_fraction"] = {"cold": traffic / (per_launch_ms * 1e-3) / 1e9 / peak,
                                         "warm": (tjw.get(dom, 0) / (per_launch_ms * 1e-3) / 1e9 / peak) if tjw.get(dom) else None,
                                         "warm_bytes_per_pair_all_passes": tjw.get("pair_total"),
                                         "alg_bytes_per_pair": 2 * ALG_BYTES_PER_MATVEC,
                                         "source": "static ncu counters (profiles/traffic.json) over this run's launch time"}
    except Exception:
        pass
    fp64_flop_per_matvec = 5.0 * (2 * n) * math.log2(2 * n)
    fp64_tflops = fp64_flop_per_matvec * ncols / (ms / args.steps * 1e-3) / 1e12
    roofline["fp64"] = {"flop_per_matvec": fp64_flop_per_matvec, "achieved_tflops": fp64_tflops, "peak_tflops": FP64_PEAK_TFLOPS,
                        "peak_source": "profiles/tools/fp64_peak.cu on this pool (round 1)", "frac": fp64_tflops / FP64_PEAK_TFLOPS}

    # ---- CPU baseline on a bounded sample of the same workload, rank 0 only
    cpu = None
    parity = parity_all
    if not args.no_cpu:
        cores = host_cores()
        sample = max(cores, 16) * 4            # ~17 CPU-seconds of work
        g.manual_seed(1234)
        Xs = np.asfortranarray(torch.randn((sample, n), dtype=torch.float64, device="cuda", generator=g).t().cpu().numpy())
        Yo, dt = cpu_reference_columns(vc, vr, Xs, cores)
        cpu = {"value": sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "first %d of the %d columns, reference algorithm (FFT length 2^21-1, pocketfft), "
                         "columns spread over %d threads, %.1f s" % (sample, ncols, cores, dt)}
        Yg = tb.mul(F, torch.from_numpy(np.ascontiguousarray(Xs.T)).cuda().t()).cpu().numpy()
        parity = max(parity, rel_err(Yg, Yo))

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "Toeplitz{Float64} n=2^20 x Matrix with %d RHS columns per GPU (BASELINE configs[1])" % ncols,
                   "n": n, "columns_per_gpu": ncols, "parallelism": "columns sharded over %d GPU(s), spectrum broadcast once" % world,
                   "spectrum_broadcast": bcast_how,
                   "l2": "inputs (8 GiB X + 8 GiB Y per GPU) far exceed the 126 MB L2; no flush needed"},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": e2e_calls * e2e_cols * n * 8,
                "d2h_bytes_per_step": e2e_calls * e2e_cols * n * 8,
                "steps": e2e_steps, "columns_per_gpu_per_step": e2e_calls * e2e_cols, "calls_per_step": e2e_calls,
                "columns_per_call": e2e_cols,
                "note": "one step = the config's %d columns per GPU through tb200_mul_host, as %d calls over the same pinned "
                        "%d-column host matrices (%.0f GiB pinned per rank); every call copies its columns host->device and "
                        "the product device->host" % (e2e_calls * e2e_cols, e2e_calls, e2e_cols, 2 * e2e_cols * n * 8 / 2**30),
                "max_abs_diff_vs_device_path": e2e_err,
                "host_gbs": 2 * e2e_val * n * 8 / 1e9,
                "host_link_ceiling_gbs": {"this_rank": link, "sum_over_ranks": link_sum,
                                          "how": "pinned 1-GiB H2D + D2H copies at once, measured after the e2e leg"},
                "numa_binding": numa, "pageable_host": pageable},
        "gpu_launches": int(launches),
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "parity_rel_err_vs_oracle": parity,
        "parity": {"max_over_ranks_rel_err_vs_oracle": parity_all, "ranks_checked": world, "columns_per_rank": len(pcols),
                   "full_batch_linearity_rel_err": lin_all, "full_batch_columns_per_rank": ncols, "tolerance": 1e-12},
        "strong_scaling": strong, "configs": configs,
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
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (and everything after it)")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1/C3/C4/C5 block")
    ap.add_argument("--quick-configs", action="store_true", help="smaller C3/C4/C5 sizes (smoke runs)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg at N > 1")
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
