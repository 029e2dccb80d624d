import torch, sys
dev="cuda"
def t(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(iters):
        s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e)*1e3)
    ts.sort(); return ts[len(ts)//2]
for mb in (184, 737, 1659):
    x=torch.empty(mb*1000*1000//4, dtype=torch.float32, device=dev)
    us=t(lambda: x.zero_()); print(f"zero_ {mb} MB: {us:.1f} us  {mb/us*1e3:.0f} GB/s write-only")
    y=torch.empty_like(x)
    us=t(lambda: y.copy_(x)); print(f"copy {mb} MB: {us:.1f} us  {2*mb/us*1e3:.0f} GB/s r+w")
    del x,y
for (B,N) in ((2,4800),(2,14400)):
    a=torch.randn(B,N,256,device=dev); b=torch.randn(B,256,N,device=dev); out=torch.empty(B,N,N,device=dev)
    for prec in ("highest","high","medium"):
        torch.set_float32_matmul_precision(prec)
        us=t(lambda: torch.bmm(a,b,out=out)); print(f"cublas bmm fp32[{prec}] B{B} N{N}: {us:.1f} us  {B*4*N*N/us/1e3:.0f} GB/s")
    ah,bh=a.half(),b.half(); oh=torch.empty(B,N,N,device=dev,dtype=torch.half)
    us=t(lambda: torch.bmm(ah,bh,out=oh)); print(f"cublas bmm fp16 out fp16 B{B} N{N}: {us:.1f} us  {B*2*N*N/us/1e3:.0f} GB/s")
