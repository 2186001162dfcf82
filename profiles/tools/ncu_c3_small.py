"""A few columns of config 3 (Circulant{ComplexF64} n = 2^22 ldiv!) for per-pass DRAM counters."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np, torch, toeplitz_b200 as tb
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    tb.set_option(k, int(v))
n, l = 1 << 22, 12
v = torch.from_numpy((0.9 ** np.arange(n)).astype(np.complex128)).cuda()
F = tb.factorize(tb.Circulant(v))
B = torch.randn((l, n), dtype=torch.complex128, device="cuda").t()
X = tb.colmajor(n, l, torch.complex128)
for _ in range(2):
    tb.circ_ldiv_(F, B, out=X)
torch.cuda.synchronize()
