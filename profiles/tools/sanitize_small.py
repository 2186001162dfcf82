"""Small shapes of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck; one tool per gpurun call):
    compute-sanitizer --tool memcheck python profiles/tools/sanitize_small.py
Every product is checked against the oracle so a sanitizer-clean run is also a correct one."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import toeplitz_b200 as tb
from oracle import toeplitz_oracle as O
from oracle import build_oracle as OC

rng = np.random.default_rng(0)
def dev(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()
def devmat(a): return torch.from_numpy(np.ascontiguousarray(a.T)).cuda().t()
def rel(a, b): return float(np.linalg.norm((np.asarray(a) - np.asarray(b)).ravel()) / max(np.linalg.norm(np.asarray(b).ravel()), 1e-300))
worst = {}
def chk(name, e, tol):
    worst[name] = e
    assert e <= tol, (name, e)

LIGHT = os.environ.get("SAN_LIGHT", "0") == "1"       # racecheck: smaller sizes
for dt, tol in ((np.float64, 1e-12), (np.complex128, 1e-12), (np.float32, 1e-5), (np.complex64, 1e-5)):
    rdt = np.zeros(1, dt).real.dtype
    for n in ((40, 300, 9000) if not LIGHT else (40, 300, 5000)):      # direct loop, one shared-memory transform, two-level
        k = np.arange(n)
        vc, vr = (0.9 ** k).astype(dt), (0.4 ** k).astype(dt)
        X = rng.standard_normal((n, 3)).astype(rdt)
        Ao, Ad = O.Toeplitz(vc, vr), tb.Toeplitz(dev(vc), dev(vr))
        chk("mul %s n=%d" % (np.dtype(dt).name, n), rel(tb.mul(Ad, devmat(X)).cpu().numpy(), O.matvec(Ao, X)), tol)
        chk("mul vec %s n=%d" % (np.dtype(dt).name, n), rel(tb.mul(Ad, dev(X[:, 0].copy())).cpu().numpy(), O.matvec(Ao, X[:, 0])), tol)
    n = 2048
    v = (0.9 ** np.arange(n)).astype(dt)
    b = rng.standard_normal(n).astype(rdt).astype(dt)
    chk("circ ldiv %s" % np.dtype(dt).name, rel(tb.ldiv(tb.Circulant(dev(v)), dev(b)).cpu().numpy(), O.circ_ldiv(O.Circulant(v), b.copy())), tol * 10)
    v101 = (0.9 ** np.arange(101)).astype(dt)
    b101 = rng.standard_normal(101).astype(rdt).astype(dt)
    chk("bluestein ldiv %s" % np.dtype(dt).name, rel(tb.ldiv(tb.Circulant(dev(v101)), dev(b101)).cpu().numpy(), O.circ_ldiv(O.Circulant(v101), b101.copy())), tol * 10)
    chk("eigvals %s" % np.dtype(dt).name, rel(tb.eigvals(tb.Circulant(dev(v))).cpu().numpy(), O.circ_eigvals(O.Circulant(v))), tol * 10)
    chk("inv %s" % np.dtype(dt).name, rel(tb.inv(tb.Circulant(dev(v))).vc.cpu().numpy(), O.circ_inv(O.Circulant(v)).vc), tol * 10)
    hv = rng.standard_normal(2 * 600 - 1).astype(rdt).astype(dt)
    hx = rng.standard_normal(600).astype(rdt).astype(dt)
    chk("hankel %s" % np.dtype(dt).name, rel(tb.mul(tb.Hankel(dev(hv), (600, 600)), dev(hx)).cpu().numpy(), O.matvec(O.Hankel(hv, size=(600, 600)), hx)), tol)

# two-level with TMA staging (plan 2^9 x 2^9)
n = 1 << 17
kk = torch.arange(n, dtype=torch.float64, device="cuda")
F = tb.factorize(tb.Toeplitz(0.9 ** kk, 0.4 ** kk))
Xd = torch.randn((3, n), dtype=torch.float64, device="cuda").t()
ref = tb.mul(F, Xd)
tb.set_option("tma", 3)
out = tb.mul(F, Xd)
tb.set_option("tma", 0)
assert torch.equal(out, ref)

# Levinson / Durbin: lane groups, one warp, generator form, multi-warp, block fallback
for n in ((40, 130, 300, 500, 600, 2100) if not LIGHT else (40, 130, 300, 500, 600)):
    ns = 9
    R = np.asfortranarray(rng.random((n - 1, ns)) / n)
    B = np.asfortranarray(rng.standard_normal((n, ns)))
    for sx in (1, 0):
        tb.set_option("lev_gen", sx)
        Xl = tb.colmajor(n, ns)
        tb.levinson_(Xl, devmat(R), devmat(B))
        chk("levinson n=%d gen=%d" % (n, sx), rel(Xl.cpu().numpy(), OC.c_levinson_batched(R, B)), 1e-12)
        Rd = np.asfortranarray(rng.random((n, ns)) / n)
        Yl = tb.colmajor(n, ns)
        tb.durbin_(Yl, devmat(Rd))
        chk("durbin n=%d gen=%d" % (n, sx), rel(Yl.cpu().numpy(), OC.c_durbin_batched(Rd)), 1e-12)
    tb.set_option("lev_gen", 1)
a = np.concatenate([[1.0], rng.random(299) / 300])
Bm = rng.standard_normal((300, 6))
chk("levinson multirhs", rel(tb.levinson(tb.SymmetricToeplitz(dev(a)), devmat(Bm)).cpu().numpy(), np.linalg.solve(O.SymmetricToeplitz(a).dense(), Bm)), 1e-10)
chk("trench", rel(tb.trench(dev(a[1:200])).cpu().numpy(), np.linalg.inv(O.SymmetricToeplitz(a[:200]).dense())), 1e-10)
chk("tri inv", rel(tb.inv(tb.LowerTriangularToeplitz(dev(0.9 ** np.arange(300)))).v.cpu().numpy(),
                   np.linalg.inv(O.LowerTriangularToeplitz(0.9 ** np.arange(300)).dense())[:, 0]), 1e-10)

# Krylov: single and batched, one-level and two-level, cgs and cg
for n, l in ((700, 5), (6000, 5)):
    v = 0.5 ** np.arange(n)
    Ao, Ad = O.SymmetricToeplitz(v), tb.SymmetricToeplitz(dev(v))
    Md = tb.factorize(tb.strang(Ad))
    B = rng.standard_normal((n, l)); B[:, 1] = 0
    tol = 100 * np.finfo(float).eps
    X = tb.colmajor(n, l); X.zero_()
    X, errs, its, flags = tb.cgs_batched(Ad, X, devmat(B), Md, 30, tol)
    chk("cgs batched n=%d" % n, rel(X.cpu().numpy(), np.linalg.solve(Ao.dense(), B)), 1e-10)
    X2 = tb.colmajor(n, l); X2.zero_()
    X2, errs, its, flags = tb.cg_batched(Ad, X2, devmat(B), Md, 30, tol)
    chk("cg batched n=%d" % n, rel(X2.cpu().numpy()[:, [0, 2, 3, 4]], np.linalg.solve(Ao.dense(), B)[:, [0, 2, 3, 4]]), 1e-10)
    x1, e1, i1, f1 = tb.cgs(Ad, torch.zeros(n, dtype=torch.float64, device="cuda"), dev(B[:, 0].copy()), Md, 30, tol)
    chk("cgs single n=%d" % n, rel(x1.cpu().numpy(), np.linalg.solve(Ao.dense(), B[:, 0])), 1e-10)
torch.cuda.synchronize()
print("sanitize_small: all checks passed; worst:", max(worst.items(), key=lambda kv: kv[1]))
