"""Single-vector products beyond the two-level plan (three-level: N' = 2^9 x 2^(l-9)) next to the largest two-level size."""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch, toeplitz_b200 as tb
from config_bench import timed
for dt, logs in ((torch.float64, (23, 24, 25, 26)), (torch.complex64, (24, 25, 26, 27))):
    for ln in logs:
        n = 1 << ln
        vc = torch.randn(n, dtype=dt, device="cuda"); vr = torch.randn(n, dtype=dt, device="cuda"); vr[0] = vc[0]
        x = torch.randn(n, dtype=dt, device="cuda"); y = torch.empty_like(x)
        F = tb.factorize(tb.Toeplitz(vc, vr))
        t = timed(lambda: tb.mul_(y, F, x), 5, warm=2)
        tf = timed(lambda: tb.factorize(tb.Toeplitz(vc, vr)), 2, warm=1)
        esz = x.element_size()
        print(json.dumps({"config": "Toeplitz{%s} n=2^%d, one vector (embedding 2^%d)" % (str(dt).split(".")[1], ln, ln + 1), "mul_ms": t * 1e3,
                          "factorize_ms": tf * 1e3, "alg_gbs": 2 * n * esz / t / 1e9}), flush=True)
        del F, vc, vr, x, y
        torch.cuda.empty_cache()
