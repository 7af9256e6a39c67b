"""Row-sliding conv micro-benchmark at the generator's 512 / 1024 px layers (batch 16, bf16, full StyledConv epilogue)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ood_gan_inversion_b200 import kernels as K  # noqa: E402


def timeit(fn, warm=3, rep=10):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(rep):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


b = int(os.environ.get('B', 16))
for ci, co, r, ys in [(32, 32, 1024, False), (64, 64, 512, True), (64, 32, 1024, True)]:
    if ci != co and r == 1024:
        continue
    x = torch.randn(b, r, r, ci, device='cuda').bfloat16()
    w = torch.randn(co, ci, 3, 3, device='cuda') * 0.1
    wp = K.pack_conv_weight(w, torch.bfloat16, False)
    d, bias, sn = torch.rand(b, co, device='cuda') + 0.5, torch.randn(co, device='cuda'), torch.rand(b, co, device='cuda') + 0.5
    noise, nw = torch.randn(b, 1, r, r, device='cuda'), torch.tensor([0.1], device='cuda')
    fn = lambda: K.conv3x3(x, wp, co, d=d, noise=noise, noise_w=nw, bias=bias, s_next=sn if ys else None, act=True, want_y=True, want_ys=ys)
    if os.environ.get('RGB'):          # the last layer with ToRGB in the epilogue: no activation written at all
        wrgb = K.torgb_weight(torch.randn(3, co, device='cuda'), torch.rand(b, co, device='cuda') + 0.5)
        rb, skip = torch.randn(3, device='cuda'), torch.randn(b, 3, r // 2, r // 2, device='cuda')
        fn = lambda: K.conv3x3(x, wp, co, d=d, noise=noise, noise_w=nw, bias=bias, act=True, want_y=False, want_ys=False,
                               rgb=(wrgb, rb, skip, K.fir_taps(gain=2.0)))
    best, med = timeit(fn)
    byt = b * r * r * (ci + co * (2 if ys else 1)) * 2 + b * r * r * 4
    fl = 2.0 * b * co * ci * 9 * r * r
    print(f'conv {ci}->{co} {r}px y{"+ys" if ys else ""}: best {best * 1e3:.1f} us median {med * 1e3:.1f} us  {byt / best / 1e6:.0f} GB/s ({byt / best / 1e6 / 6534.8:.3f} of HBM)  {fl / best / 1e9:.0f} TFLOP/s')
