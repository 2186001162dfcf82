"""Timed comparison ONLY (never on the product path): the same batched Toeplitz product done with cuFFT through
torch.fft (rfft of the zero-padded real columns, pointwise multiply by the half spectrum, irfft), chunked so the
padded intermediates fit, against libtoeplitz_b200 on identical inputs.  C2: n = 2^20, Float64, 1024 columns."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import toeplitz_b200 as tb


def timed(fn, reps=3, warm=2):
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


def main():
    n, cols = 1 << 20, 1024
    k = np.arange(n, dtype=np.float64)
    vc, vr = torch.from_numpy(0.9 ** k).cuda(), torch.from_numpy(0.4 ** k).cuda()
    X = torch.randn((cols, n), dtype=torch.float64, device="cuda")          # row-major rows = columns of the Julia matrix
    Y = torch.empty_like(X)
    # cuFFT path: embedding of length 2n, real-to-complex transforms
    c = torch.zeros(2 * n, dtype=torch.float64, device="cuda")
    c[:n] = vc
    c[n + 1:] = vr[1:].flip(0)
    spec = torch.fft.rfft(c)
    chunk = 64

    def cufft_step():
        for j in range(0, cols, chunk):
            Xf = torch.fft.rfft(X[j:j + chunk], n=2 * n, dim=1)
            Xf *= spec
            Y[j:j + chunk] = torch.fft.irfft(Xf, n=2 * n, dim=1)[:, :n]

    t_cufft = timed(cufft_step)
    Yref = Y[:4].clone()
    F = tb.factorize(tb.Toeplitz(vc, vr))
    Xc, Yc = X.t(), torch.empty_like(X).t()
    t_ours = timed(lambda: tb.mul_(Yc, F, Xc), reps=5, warm=3)
    err = float(torch.linalg.vector_norm(Yc[:, :4].t() - Yref) / torch.linalg.vector_norm(Yref))
    print(json.dumps({"config": "C2 Toeplitz{Float64} n=2^20 x 1024 columns", "cufft_rfft_pipeline_matvecs_per_s": cols / t_cufft,
                      "cufft_ms_per_batch": t_cufft * 1e3, "toeplitz_b200_matvecs_per_s": cols / t_ours, "ours_ms_per_batch": t_ours * 1e3,
                      "speedup_vs_cufft_pipeline": t_cufft / t_ours, "rel_diff_ours_vs_cufft": err,
                      "note": "cuFFT via torch.fft (R2C/C2R, 64-column chunks, separate pointwise multiply and slice copy)"}))


def c3():
    """Circulant{ComplexF64} n = 2^22 ldiv! over 256 right-hand sides: cuFFT Z2Z forward, pointwise divide, inverse."""
    n, cols = 1 << 22, 256
    v = torch.from_numpy((0.9 ** np.arange(n)).astype(np.complex128)).cuda()
    B = torch.randn((cols, n), dtype=torch.complex128, device="cuda")
    Xo = torch.empty_like(B)
    ispec = 1.0 / torch.fft.fft(v)
    chunk = 16

    def cufft_step():
        for j in range(0, cols, chunk):
            Bf = torch.fft.fft(B[j:j + chunk], dim=1)
            Bf *= ispec
            Xo[j:j + chunk] = torch.fft.ifft(Bf, dim=1)

    t_cufft = timed(cufft_step)
    ref = Xo[:2].clone()
    F = tb.factorize(tb.Circulant(v))
    Bc, Xc = B.t(), torch.empty_like(B).t()
    t_ours = timed(lambda: tb.circ_ldiv_(F, Bc, out=Xc), reps=3, warm=2)
    err = float(torch.linalg.vector_norm(Xc[:, :2].t() - ref) / torch.linalg.vector_norm(ref))
    print(json.dumps({"config": "C3 Circulant{ComplexF64} n=2^22 ldiv!, 256 RHS", "cufft_pipeline_solves_per_s": cols / t_cufft,
                      "cufft_ms_per_batch": t_cufft * 1e3, "toeplitz_b200_solves_per_s": cols / t_ours, "ours_ms_per_batch": t_ours * 1e3,
                      "speedup_vs_cufft_pipeline": t_cufft / t_ours, "rel_diff_ours_vs_cufft": err,
                      "note": "cuFFT via torch.fft (Z2Z, 16-column chunks, separate pointwise multiply)"}))


def c5a():
    """Hankel{ComplexF32} n = 2^24 matvec: cuFFT C2C of the zero-padded reversed x (2^25), pointwise multiply, inverse."""
    n = 1 << 24
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    v = torch.randn(2 * n - 1, dtype=torch.complex64, device="cuda", generator=g)
    x = torch.randn(n, dtype=torch.complex64, device="cuda", generator=g)
    # H x = T reverse(x) with T = Toeplitz(vc = v[n-1:], vr = reverse(v[:n])); embedding c = [vc; 0; vr[n-1..1]]
    c = torch.zeros(2 * n, dtype=torch.complex64, device="cuda")
    c[:n] = v[n - 1:]
    c[n + 1:] = v[:n - 1]
    spec = torch.fft.fft(c)
    y = torch.empty_like(x)
    pad = torch.zeros(2 * n, dtype=torch.complex64, device="cuda")

    def cufft_step():
        pad[:n] = x.flip(0)
        f = torch.fft.fft(pad)
        f *= spec
        y.copy_(torch.fft.ifft(f)[:n])

    t_cufft = timed(cufft_step, reps=10)
    ref = y.clone()
    F = tb.factorize(tb.Hankel(v, (n, n)))
    yo = torch.empty_like(x)
    t_ours = timed(lambda: tb.mul_(yo, F, x), reps=20, warm=3)
    err = float(torch.linalg.vector_norm(yo - ref) / torch.linalg.vector_norm(ref))
    print(json.dumps({"config": "C5a Hankel{ComplexF32} n=2^24 matvec", "cufft_pipeline_ms": t_cufft * 1e3, "ours_ms": t_ours * 1e3,
                      "speedup_vs_cufft_pipeline": t_cufft / t_ours, "rel_diff_ours_vs_cufft": err,
                      "note": "cuFFT via torch.fft (C2C 2^25, separate pad / reverse / multiply / slice kernels)"}))


if __name__ == "__main__":
    which = sys.argv[1:] or ["c2"]
    for w in which:
        {"c2": main, "c3": c3, "c5a": c5a}[w]()
        torch.cuda.empty_cache()
