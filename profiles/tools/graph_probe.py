"""Is the headline step bound by launch overhead?  Capture one tb.mul_ of the C2 batch (1536 kernel launches over two
helper streams) into a CUDA graph and compare replay time with the eager call."""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch, toeplitz_b200 as tb
n, cols = 1 << 20, 1024
torch.manual_seed(0)
vc = torch.randn(n, dtype=torch.float64, device="cuda"); vr = torch.randn(n, dtype=torch.float64, device="cuda"); vr[0] = vc[0]
F = tb.factorize(tb.Toeplitz(vc, vr))
X = tb.colmajor(n, cols); X.copy_(torch.randn(n, cols, dtype=torch.float64, device="cuda"))
Y = tb.colmajor(n, cols)
def timed(fn, reps=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
eager = timed(lambda: tb.mul_(Y, F, X))
Yref = Y.clone()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    tb.mul_(Y, F, X)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        tb.mul_(Y, F, X)
Y.zero_()
replay = timed(lambda: g.replay())
print(json.dumps({"eager_ms": eager, "graph_replay_ms": replay, "same": bool(torch.equal(Y, Yref))}))
