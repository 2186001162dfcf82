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


if __name__ == "__main__":
    main()
