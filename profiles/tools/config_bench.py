"""Timings of the non-headline BASELINE configs (C1, C3, C4, C5) on one B200 -> JSON lines.
CUDA events on the launching stream, 3 warm-ups, inputs resident in HBM.  Not the driver's bench
(that is bench.py, config 2); the numbers are quoted in DESIGN.md."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import toeplitz_b200 as tb

HBM = 6553.0


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def out(**kw):
    print(json.dumps(kw), flush=True)


def c1():
    n = 1 << 16
    k = np.arange(n, dtype=np.float64)
    A = tb.Toeplitz(torch.from_numpy(0.9 ** k).cuda(), torch.from_numpy(0.4 ** k).cuda())
    x = torch.randn(n, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    F = tb.factorize(A)
    t = timed(lambda: tb.mul_(y, F, x), 200)
    tf = timed(lambda: tb.factorize(A), 50)
    t_ref = timed(lambda: tb.mul_(y, A, x), 50)          # reference semantics: re-factorize every call
    out(config="C1 Toeplitz{Float64} n=2^16 single vector", mul_us=t * 1e6, factorize_us=tf * 1e6,
        mul_with_refactorize_us=t_ref * 1e6, alg_gbs=2 * n * 8 / t / 1e9)


def c3():
    n, l = 1 << 22, 256
    v = torch.from_numpy((0.9 ** np.arange(n)).astype(np.complex128)).cuda()
    C = tb.Circulant(v)
    tf = timed(lambda: tb.factorize(C), 20)
    F = tb.factorize(C)
    B = torch.randn((l, n), dtype=torch.complex128, device="cuda").t()
    t = timed(lambda: tb.ldiv_(F, B), 3, warm=2)
    te = timed(lambda: tb.eigvals(F), 20)
    out(config="C3 Circulant{ComplexF64} n=2^22, 256 RHS", ldiv_solves_per_s=l / t, ldiv_ms_per_batch=t * 1e3,
        factorize_ms=tf * 1e3, eigvals_ms=te * 1e3, alg_gbs=l * 2 * n * 16 / t / 1e9, frac_hbm=l * 2 * n * 16 / t / 1e9 / HBM)


def c4():
    n, nsys = 512, 1000000
    R = (torch.rand((nsys, n), dtype=torch.float64, device="cuda") / n).t()
    B = torch.randn((nsys, n), dtype=torch.float64, device="cuda").t()
    X = tb.colmajor(n, nsys)
    Rl = R[: n - 1, :]
    t = timed(lambda: tb.levinson_(X, Rl, B), 3, warm=1)
    Y = tb.colmajor(n, nsys)
    td = timed(lambda: tb.durbin_(Y, R), 3, warm=1)
    out(config="C4 SymmetricToeplitz{Float64} n=512 x 10^6 systems", levinson_solves_per_s=nsys / t, durbin_solves_per_s=nsys / td,
        levinson_tflops=4.0 * n * n * nsys / t / 1e12, durbin_tflops=2.0 * n * n * nsys / td / 1e12,
        levinson_alg_gbs=nsys * 12280 / t / 1e9, durbin_alg_gbs=nsys * 8192 / td / 1e9)


def c5():
    n = 1 << 24
    v = torch.randn(2 * n - 1, dtype=torch.complex64, device="cuda")
    x = torch.randn(n, dtype=torch.complex64, device="cuda")
    y = torch.empty_like(x)
    H = tb.Hankel(v, (n, n))
    F = tb.factorize(H)
    t = timed(lambda: tb.mul_(y, F, x), 20)
    out(config="C5a Hankel{ComplexF32} n=2^24 matvec", matvecs_per_s=1 / t, ms=t * 1e3, alg_gbs=2 * n * 8 / t / 1e9,
        frac_hbm=2 * n * 8 / t / 1e9 / HBM)
    del F, H, v
    A = tb.SymmetricToeplitz(torch.from_numpy((0.5 ** np.arange(n)).astype(np.complex64)).cuda())
    FA = tb.factorize(A)
    M = tb.factorize(tb.strang(A))
    b = torch.randn(n, dtype=torch.complex64, device="cuda")
    its = 20
    torch.cuda.synchronize()
    tb.cgs(FA, torch.zeros_like(b), b, M, 2, 0.0)
    t0 = time.perf_counter()
    _, err, it, flag = tb.cgs(FA, torch.zeros_like(b), b, M, its, 0.0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out(config="C5b SymmetricToeplitz{ComplexF32} n=2^24 cgs + Strang", iterations_per_s=it / dt, ms_per_iteration=dt / it * 1e3,
        iterations=it, flag=flag, final_rel_residual=err, alg_gbs=19 * n * 8 * it / dt / 1e9)


def c4m():
    """one matrix, many right-hand sides (shared y-recursion) vs the replicated batched kernel"""
    n, nrhs = 512, 200000
    a = torch.cat([torch.ones(1, dtype=torch.float64, device="cuda"), torch.rand(n - 1, dtype=torch.float64, device="cuda") / n])
    T = tb.SymmetricToeplitz(a)
    B = torch.randn((nrhs, n), dtype=torch.float64, device="cuda").t()
    t = timed(lambda: tb.levinson(T, B), 3, warm=1)
    R = a[1:].unsqueeze(1).expand(-1, nrhs)
    R = tb.to_colmajor(R)
    X = tb.colmajor(n, nrhs)
    t2 = timed(lambda: tb.levinson_(X, R, B), 3, warm=1)
    out(config="C4m SymmetricToeplitz{Float64} n=512, one matrix x 2e5 right-hand sides", shared_recursion_solves_per_s=nrhs / t,
        replicated_solves_per_s=nrhs / t2)


def c4s():
    """Levinson / Durbin at smaller orders (same kernels as C4)"""
    for n in (32, 64, 128, 256, 768, 1024, 2048):
        nsys = 1 << 20 if n <= 256 else 1 << 17
        R = (torch.rand((nsys, n), dtype=torch.float64, device="cuda") / n).t()
        B = torch.randn((nsys, n), dtype=torch.float64, device="cuda").t()
        X = tb.colmajor(n, nsys)
        Rl = R[: n - 1, :]
        t = timed(lambda: tb.levinson_(X, Rl, B), 3, warm=1)
        Y = tb.colmajor(n, nsys)
        td = timed(lambda: tb.durbin_(Y, R), 3, warm=1)
        out(config="C4s n=%d x %d systems" % (n, nsys), levinson_solves_per_s=nsys / t, durbin_solves_per_s=nsys / td,
            levinson_tflops=4.0 * n * n * nsys / t / 1e12, durbin_tflops=2.0 * n * n * nsys / td / 1e12,
            levinson_alg_gbs=nsys * (3 * n * 8) / t / 1e9)


def sweep():
    """Batched Toeplitz matvec across sizes: algorithmic GB/s (read x + write y) at ~512 MiB per operand.
    SWEEP_DTYPE = f64 (default) | f32 | c64 | c128."""
    name = os.environ.get("SWEEP_DTYPE", "f64")
    dt = {"f64": torch.float64, "f32": torch.float32, "c64": torch.complex64, "c128": torch.complex128}[name]
    esz = torch.empty(0, dtype=dt).element_size()
    for ln in range(int(os.environ.get("SWEEP_LO", 8)), int(os.environ.get("SWEEP_HI", 24))):
        n = 1 << ln
        cols = max(2, min(1 << 20, (512 << 20) // (esz * n)))
        k = np.arange(n, dtype=np.float64)
        A = tb.Toeplitz(torch.from_numpy(0.9 ** k).cuda().to(dt), torch.from_numpy(0.4 ** k).cuda().to(dt))
        F = tb.factorize(A)
        X = tb.colmajor(n, cols, dt)
        X.copy_(torch.randn(n, cols, dtype=dt, device="cuda"))
        Y = tb.colmajor(n, cols, dt)
        t = timed(lambda: tb.mul_(Y, F, X), 5)
        out(config="sweep Toeplitz{%s} n=2^%d x %d columns" % (name, ln, cols), matvecs_per_s=cols / t,
            alg_gbs=2 * esz * n * cols / t / 1e9, frac_hbm=2 * esz * n * cols / t / 1e9 / HBM)
        del X, Y, F


def c6(seeds=(0,)):
    """Toeplitz \\ B with many right-hand sides: batched cgs + Strang vs the reference's column loop of single solves."""
    for n, l, seed in [(1 << 16, 256, sd) for sd in seeds] + [(1 << 20, 64, 0)]:
        k = np.arange(n, dtype=np.float64)
        A = tb.Toeplitz(torch.from_numpy(0.9 ** k).cuda(), torch.from_numpy(0.4 ** k).cuda())
        FA, M = tb.factorize(A), tb.factorize(tb.strang(A))
        B = tb.colmajor(n, l)
        gen = torch.Generator(device="cuda")
        gen.manual_seed(seed)
        B.copy_(torch.randn(n, l, dtype=torch.float64, device="cuda", generator=gen))
        tol = 100 * 2.220446049250313e-16
        X = tb.colmajor(n, l)

        def batched():
            X.zero_()
            return tb.cgs_batched(FA, X, B, M, 1000, tol)

        def looped(cols=8):
            for j in range(cols):
                tb.cgs(FA, torch.zeros(n, dtype=torch.float64, device="cuda"), B[:, j].contiguous(), M, 1000, tol)

        _, errs, its, flags = batched()
        torch.cuda.synchronize()
        t0 = time.perf_counter(); batched(); torch.cuda.synchronize(); tb_ = time.perf_counter() - t0
        looped(2)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); looped(8); torch.cuda.synchronize(); tl = (time.perf_counter() - t0) / 8
        out(config="C6 Toeplitz{Float64} n=2^%d \\ %d right-hand sides, cgs + Strang, seed %d" % (int(np.log2(n)), l, seed),
            iteration_histogram={int(k): int(v) for k, v in zip(*np.unique(np.minimum(its, 1000), return_counts=True))},
            batched_solves_per_s=l / tb_, column_loop_solves_per_s=1.0 / tl, iterations=[min(its), max(its)],
            flags=sorted(set(flags)), max_rel_residual=max(errs))


def c6scan():
    """look for right-hand-side draws with a straggler column (one column running to max_it while the rest stop at 3-6)"""
    c6(seeds=tuple(range(12)))


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if "=" not in a]
    for kv in (a for a in sys.argv[1:] if "=" in a):       # tuning: library options, e.g. l2max_f64=13
        k, v = kv.split("=")
        tb.set_option(k, int(v))
    for w in (args or ["c1", "c3", "c4", "c5"]):
        globals()[w]()
