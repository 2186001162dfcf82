import sys, torch, numpy as np, time
sys.path.insert(0,'/root/repo')
import toeplitz_b200 as tb
n=2048
vc=torch.randn(n,dtype=torch.complex128,device='cuda'); vr=torch.randn(n,dtype=torch.complex128,device='cuda'); vr[0]=vc[0]
F=tb.factorize(tb.Toeplitz(vc,vr))
for cols in (296, 592, 2368, 16384, 65536):
    X=torch.randn((cols,n),dtype=torch.complex128,device='cuda').t(); Y=tb.colmajor(n,cols,torch.complex128)
    for _ in range(3): tb.mul_(Y,F,X)
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    reps=5
    e0.record()
    for _ in range(reps): tb.mul_(Y,F,X)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/reps
    print("cols %6d: %.3f ms  -> %.2f us per row per SM (148 SMs), %.1f GB/s in+out" % (cols, ms, ms*1e3/(cols/148), cols*n*16*2/ms/1e6))
