"""Batched cgs + Strang at n = 2^16 x 256 columns (config 6 regime: launch-bound iterations) for a launch list."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np, torch, toeplitz_b200 as tb
n, l = 1 << 16, 256
k = np.arange(n, dtype=np.float64)
A = tb.Toeplitz(torch.from_numpy(0.9 ** k).cuda(), torch.from_numpy(0.4 ** k).cuda())
FA, M = tb.factorize(A), tb.factorize(tb.strang(A))
B = tb.colmajor(n, l); B.copy_(torch.randn(n, l, dtype=torch.float64, device="cuda"))
for _ in range(2):
    X = tb.colmajor(n, l); X.zero_()
    tb.cgs_batched(FA, X, B, M, 1000, 100 * 2.220446049250313e-16)
torch.cuda.synchronize()
