"""Two Hankel{ComplexF32} n = 2^24 products (config 5a) for `ncu --set full`."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch, toeplitz_b200 as tb
n = 1 << 24
v = torch.randn(2 * n - 1, dtype=torch.complex64, device="cuda")
x = torch.randn(n, dtype=torch.complex64, device="cuda")
y = torch.empty_like(x)
F = tb.factorize(tb.Hankel(v, (n, n)))
for _ in range(2):
    tb.mul_(y, F, x)
torch.cuda.synchronize()
