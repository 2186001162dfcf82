import torch,time
torch.cuda.init()
for mb in (4,8,16,32,64,128,256,1024):
    n=mb*(1<<20)//8
    a=torch.randn(n,dtype=torch.float64,device='cuda'); b=torch.empty_like(a)
    for _ in range(5): b.copy_(a)
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    reps=200 if mb<=64 else 20
    e0.record()
    for _ in range(reps): b.copy_(a)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/reps
    print("copy %4d MiB -> %4d MiB: %7.1f us, %6.0f GB/s (r+w)"%(mb,mb,ms*1e3, 2*mb*(1<<20)/ms/1e6))
# read-only reduction
for mb in (16,32,64,1024):
    n=mb*(1<<20)//8
    a=torch.randn(n,dtype=torch.float64,device='cuda')
    for _ in range(5): a.sum()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    reps=100
    e0.record()
    for _ in range(reps): a.sum()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/reps
    print("sum  %4d MiB: %7.1f us, %6.0f GB/s (read)"%(mb,ms*1e3, mb*(1<<20)/ms/1e6))
