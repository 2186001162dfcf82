"""A few column pairs of the headline config (for `ncu --set full`, which replays every kernel ~40 times)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch, toeplitz_b200 as tb
n, cols = 1 << 20, 8
torch.manual_seed(0)
vc = torch.randn(n, dtype=torch.float64, device="cuda"); vr = torch.randn(n, dtype=torch.float64, device="cuda"); vr[0] = vc[0]
F = tb.factorize(tb.Toeplitz(vc, vr))
X = tb.colmajor(n, cols, torch.float64, "cuda"); X.copy_(torch.randn(n, cols, dtype=torch.float64, device="cuda"))
Y = tb.colmajor(n, cols, torch.float64, "cuda")
for _ in range(2):
    tb.mul_(Y, F, X)
torch.cuda.synchronize()
