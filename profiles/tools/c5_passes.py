"""Per-pass times of config 5a (Hankel{ComplexF32} n = 2^24) under plan options (column tile width, level split)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch, toeplitz_b200 as tb
n = 1 << 24
v = torch.randn(2 * n - 1, dtype=torch.complex64, device="cuda")
x = torch.randn(n, dtype=torch.complex64, device="cuda")
y = torch.empty_like(x)
for opt in ({}, {"ta_cap": 4}, {"ta_cap": 0, "force_l1": 12}, {"force_l1": 0}):
    for k, val in opt.items():
        tb.set_option(k, val)
    F = tb.factorize(tb.Hankel(v, (n, n)))
    for _ in range(3):
        tb.mul_(y, F, x)
    torch.cuda.synchronize()
    tb.set_option("time_passes", 1); tb.pass_times()
    for _ in range(10):
        tb.mul_(y, F, x)
    torch.cuda.synchronize()
    ms, cnt = tb.pass_times()
    tb.set_option("time_passes", 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        tb.mul_(y, F, x)
    e1.record(); torch.cuda.synchronize()
    print("total ms", round(e0.elapsed_time(e1) / 10, 4), opt, {k: round(1e3 * ms[k] / max(cnt[k], 1), 1) for k in ms if cnt[k]}, flush=True)
