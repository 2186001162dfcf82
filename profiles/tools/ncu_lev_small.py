"""A few waves of the C4 Levinson / Durbin kernels (for `ncu --set full`)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch, toeplitz_b200 as tb
n, nsys = 512, 148 * 16 * 4
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    tb.set_option(k, int(v))
R = (torch.rand((nsys, n), dtype=torch.float64, device="cuda") / n).t()
B = torch.randn((nsys, n), dtype=torch.float64, device="cuda").t()
X = tb.colmajor(n, nsys)
Y = tb.colmajor(n, nsys)
for _ in range(2):
    tb.levinson_(X, R[: n - 1, :], B)
    tb.durbin_(Y, R)
torch.cuda.synchronize()
