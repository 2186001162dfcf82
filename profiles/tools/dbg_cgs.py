"""Debug: reported vs true residual of cgs + Strang on KMS systems (device products and sampled dense rows)."""
import math, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import toeplitz_b200 as tb

def run(log2n, dtype, tol, max_it):
    n = 1 << log2n
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    npd = {torch.complex64: np.complex64, torch.float64: np.float64, torch.complex128: np.complex128, torch.float32: np.float32}[dtype]
    vck = (0.5 ** np.arange(n)).astype(npd)
    A = tb.SymmetricToeplitz(torch.from_numpy(vck).cuda())
    FA = tb.factorize(A); M = tb.factorize(tb.strang(A))
    b = torch.randn(n, dtype=dtype, device="cuda", generator=g)
    x, err, it, flag = tb.cgs(FA, torch.zeros_like(b), b, M, max_it, tol)
    r = b - tb.mul(FA, x)
    true = float(torch.linalg.norm(r) / torch.linalg.norm(b))
    xs, bh = x.cpu().numpy().astype(np.complex128), b.cpu().numpy().astype(np.complex128)
    col = 0.5 ** np.arange(64)
    res = {}
    for i in (0, 5, n // 2, n - 6, n - 1):
        lo, hi = max(0, i - 63), min(n, i + 64)
        res[i] = abs(bh[i] - np.dot(col[np.abs(np.arange(lo, hi) - i)], xs[lo:hi]))
    rd = r.cpu().numpy()
    print("n=2^%d %s: it=%d flag=%d reported=%.3e true(device mul)=%.3e rms|b|=%.3f dense rows=%s device r at rows=%s" % (
        log2n, str(dtype).split(".")[-1], it, flag, err, true, float(torch.linalg.norm(b)) / math.sqrt(n),
        {k: "%.2e" % v for k, v in res.items()}, {k: "%.2e" % abs(rd[k]) for k in res}), flush=True)

for l in (12, 16, 20, 24):
    run(l, torch.complex64, 1e-5, 1000)
run(20, torch.float64, 100 * 2.2e-16, 1000)
run(20, torch.complex128, 100 * 2.2e-16, 1000)
run(24, torch.complex64, 0.0, 20)
