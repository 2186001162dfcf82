"""Phase timeline of pass B (k_rows) from per-warp clock64() marks (TB_EXP_TRACE build).
marks: 0 start, 1 row loaded, 2 forward done, 3 spectrum applied, 4 inverse done, 5 stored."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import toeplitz_b200 as tb
from toeplitz_b200 import _lib
L = _lib.lib()
L.tb200_debug_set_trace.argtypes = [ctypes.c_void_p]
n = 1 << 20
k = np.arange(n, dtype=np.float64)
F = tb.factorize(tb.Toeplitz(torch.from_numpy(0.9 ** k).cuda(), torch.from_numpy(0.4 ** k).cuda()))
X = torch.randn((8, n), dtype=torch.float64, device="cuda").t()
Y = tb.colmajor(n, 8)
tb.set_option("overlap_streams", 1)
for _ in range(3):
    tb.mul_(Y, F, X)
trace = torch.zeros((512 * 8, 8), dtype=torch.int64, device="cuda")
L.tb200_debug_set_trace(ctypes.c_void_p(trace.data_ptr()))
Xp = X[:, :2].t().contiguous().t()
Yp = tb.colmajor(n, 2)
tb.mul_(Yp, F, Xp)          # one pair: one pass-B launch of 512 CTAs x 8 warps
torch.cuda.synchronize()
L.tb200_debug_set_trace(None)
tr = trace.cpu().numpy().reshape(512, 8, 8)
t0 = tr[:, :, 0].min()
d = np.diff(tr[:, :, :6], axis=2) / 1.965e3     # us per phase per warp
names = ["load wait", "forward", "spectrum", "inverse", "store"]
print("per-warp phase durations (us): mean / p10 / p90 over %d warps" % (512 * 8))
for i, nm in enumerate(names):
    v = d[:, :, i].ravel()
    print("  %-10s %6.2f  %6.2f  %6.2f" % (nm, v.mean(), np.percentile(v, 10), np.percentile(v, 90)))
cta_start = (tr[:, :, 0].min(axis=1) - t0) / 1.965e3
cta_end = (tr[:, :, 5].max(axis=1) - t0) / 1.965e3
print("CTA duration us: mean %.2f  min %.2f  max %.2f;  kernel span %.2f us" % ((cta_end - cta_start).mean(), (cta_end - cta_start).min(), (cta_end - cta_start).max(), cta_end.max()))
smid = tr[:, 0, 7]
for sm in (0, 1, 77):
    idx = np.where(smid == sm)[0]
    print("SM %d CTAs:" % sm, [(int(i), round(float(cta_start[i]), 2), round(float(cta_end[i]), 2)) for i in idx])
    for i in idx:
        w = tr[i, 0, :6]
        print("    cta %d warp0 marks (us):" % i, [round(float((x - t0) / 1.965e3), 2) for x in w])
