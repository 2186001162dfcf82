"""Three cgs + Strang iterations of config 5b (SymmetricToeplitz{ComplexF32} n = 2^24) for a per-kernel launch list."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np, torch, toeplitz_b200 as tb
n = 1 << 24
A = tb.SymmetricToeplitz(torch.from_numpy((0.5 ** np.arange(n)).astype(np.complex64)).cuda())
FA = tb.factorize(A)
M = tb.factorize(tb.strang(A))
b = torch.randn(n, dtype=torch.complex64, device="cuda")
tb.cgs(FA, torch.zeros_like(b), b, M, 2, 0.0)
torch.cuda.synchronize()
tb.cgs(FA, torch.zeros_like(b), b, M, 3, 0.0)
torch.cuda.synchronize()
