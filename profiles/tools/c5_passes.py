import sys, os
sys.path.insert(0, "/root/repo")
import torch, toeplitz_b200 as tb
n = 1 << 24
v = torch.randn(2 * n - 1, dtype=torch.complex64, device="cuda")
x = torch.randn(n, dtype=torch.complex64, device="cuda")
y = torch.empty_like(x)
F = tb.factorize(tb.Hankel(v, (n, n)))
for opt in ({}, {"force_l1": 12}, {"force_l1": 12, "pipe_b": 1}, {"force_l1": 12, "pipe_b": 1, "ta_cap": 4}):
    for k, val in opt.items():
        tb.set_option(k, val)
    if opt:
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
    print("total ms", round(e0.elapsed_time(e1) / 10, 4), end=" ")
    print(opt, {k: round(1e3 * ms[k] / max(cnt[k], 1), 1) for k in ms if cnt[k]}, flush=True)
