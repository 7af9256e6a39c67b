import torch, sys, os
sys.path.insert(0, '/root/repo')
from ood_gan_inversion_b200 import kernels as K
def timeit(fn, warm=3, rep=10):
    for _ in range(warm): fn()
    best=1e9
    for _ in range(rep):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
    return best
taps=K.fir_taps(gain=2.0)
for (b,c,r) in [(16,32,1024),(16,64,512),(16,128,256),(16,256,128),(16,512,64)]:
    t=torch.randn(b,r+1,r+1,c,device='cuda').bfloat16()
    d=torch.rand(b,c,device='cuda')+0.5; s=torch.rand(b,c,device='cuda')+0.5
    noise=torch.randn(b,1,r,r,device='cuda'); nw=torch.tensor([0.1],device='cuda'); bias=torch.randn(c,device='cuda')
    ms=timeit(lambda: K.blur_act(t,taps,d=d,noise=noise,noise_w=nw,bias=bias,s_next=s,act=True,want_y=False,want_ys=True))
    byt=b*c*((r+1)**2+r*r)*2+b*r*r*4
    print(f'blur_act bf16 B{b} C{c} R{r}: {ms*1e3:.1f} us  {byt/ms/1e6:.0f} GB/s  frac {byt/ms/1e6/6534.8:.3f}')
