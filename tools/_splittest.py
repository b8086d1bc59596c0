import torch, sys, time
sys.path.insert(0,'/root/repo')
from e2e_asr_pytorch_b200.stepper import _split3
dev='cuda'
torch.backends.cuda.matmul.allow_tf32=False
torch.backends.cuda.matmul.allow_bf16_reduced_precision_reduction=False
g=torch.Generator().manual_seed(0)
def bench(f, it=20):
    for _ in range(3): f()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): y=f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/it
for n,k,m in [(8192,2048,4096),(256,2048,4096),(8192,1240,1200)]:
    x=(torch.randn(n,k,generator=g)*3).to(dev); w=(torch.randn(m,k,generator=g)/k**0.5).to(dev); b=torch.randn(m,generator=g).to(dev)
    want=x.double()@w.double().t()+b.double()
    w1,w2,w3=_split3(w)
    B0=w1.t().contiguous(); B1=torch.cat([w2,w1],1).t().contiguous(); B2=torch.cat([w3,w2,w1],1).t().contiguous()
    def split3class():
        a1,a2,a3=_split3(x)
        c2=torch.mm(torch.cat([a1,a2,a3],1),B2,out_dtype=torch.float32)
        c1=torch.mm(torch.cat([a1,a2],1),B1,out_dtype=torch.float32)
        c0=torch.mm(a1,B0,out_dtype=torch.float32)
        return ((c2+c1)+c0)+b
    def split3class_addmm():
        a1,a2,a3=_split3(x)
        c=torch.mm(torch.cat([a1,a2,a3],1),B2,out_dtype=torch.float32)
        c=torch.addmm(c,torch.cat([a1,a2],1),B1,out_dtype=torch.float32)
        c=torch.addmm(c,a1,B0,out_dtype=torch.float32)
        return c+b
    def fp32(): return torch.nn.functional.linear(x,w,b)
    print(n,k,m,"fp32 err %.3g  %.3f ms"%((fp32().double()-want).abs().max().item(), bench(fp32)))
    print("   3-class err %.3g  %.3f ms"%((split3class().double()-want).abs().max().item(), bench(split3class)))
    try:
        print("   3-class addmm err %.3g  %.3f ms"%((split3class_addmm().double()-want).abs().max().item(), bench(split3class_addmm)))
    except Exception as e: print("   addmm variant failed:", str(e)[:150])
