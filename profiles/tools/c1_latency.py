"""Config 1 (Toeplitz{Float64} n = 2^16, one vector): where do the ~18 us per `mul!` go?  Back-to-back calls through the
Python mirror, the same through a captured CUDA graph (device time of the three dependent launches without host
launch cost), and the per-pass device times."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import toeplitz_b200 as tb


def ev_time(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    out = []
    for log2n in [int(a) for a in sys.argv[1:]] or (12, 14, 16, 18, 20):
        n = 1 << log2n
        k = np.arange(n, dtype=np.float64)
        F = tb.factorize(tb.Toeplitz(torch.from_numpy(0.9 ** k).cuda(), torch.from_numpy(0.4 ** k).cuda()))
        x = torch.randn(n, dtype=torch.float64, device="cuda")
        y = torch.empty_like(x)
        for _ in range(5):
            tb.mul_(y, F, x)
        eager = ev_time(lambda: tb.mul_(y, F, x), 300)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(300):
            tb.mul_(y, F, x)
        host = (time.perf_counter() - t0) / 300 * 1e6
        torch.cuda.synchronize()
        # the C ABI alone (what a ccall from the Julia extension pays): arguments prepared once
        from toeplitz_b200 import _lib as L
        lib, one, zero = L.lib(), L.dbl2(1.0), L.dbl2(0.0)
        st = torch.cuda.current_stream().cuda_stream
        dt = 1                                    # TB200 Float64 code, see include/toeplitz_b200.h
        from toeplitz_b200.api import _DT
        dt = _DT[torch.float64]
        xp, yp = x.data_ptr(), y.data_ptr()
        call = lambda: lib.tb200_mul(F._h, yp, n, dt, xp, n, dt, 1, one, zero, st)
        for _ in range(10):
            call()
        c_eager = ev_time(call, 300)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(300):
            call()
        c_host = (time.perf_counter() - t0) / 300 * 1e6
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s):
            for _ in range(3):
                tb.mul_(y, F, x)
            s.synchronize()
            with torch.cuda.graph(g, stream=s):
                tb.mul_(y, F, x)
        for _ in range(5):
            g.replay()
        graph = ev_time(g.replay, 300)
        g10 = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g10, stream=s):
                for _ in range(10):
                    tb.mul_(y, F, x)
        g10.replay()
        graph10 = ev_time(g10.replay, 50) / 10
        tb.set_option("time_passes", 1)
        tb.pass_times()
        for _ in range(50):
            tb.mul_(y, F, x)
        torch.cuda.synchronize()
        pt = tb.pass_times()
        tb.set_option("time_passes", 0)
        rec = {"n": n, "eager_us": eager, "host_enqueue_us": host, "c_abi_eager_us": c_eager, "c_abi_host_enqueue_us": c_host, "graph_replay_us": graph, "graph_of_10_per_mul_us": graph10,
               "pass_times": pt}
        print(json.dumps(rec), flush=True)
        out.append(rec)


if __name__ == "__main__":
    main()
