"""A/B of the CTA-pair (tcgen05 cta_group::2) form of conv_tc on the step's large convolutions (batch 16): OOD_CTA2 = 0 / 1 / 2."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ood_gan_inversion_b200 import kernels as K  # noqa: E402

dev = 'cuda'
B = 16


def timeit(fn, n=12):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


cases = [(256, 256, 256, 'plain'), (256, 256, 256, 'stats'), (512, 512, 128, 'stats'), (1024, 1024, 64, 'stats'), (256, 512, 128, 'seed'),
         (512, 1024, 64, 'seed'), (128, 128, 256, 'plain'), (256, 256, 128, 'plain'), (512, 512, 64, 'plain')]
for ci, co, r, kind in cases:
    g = torch.Generator(device=dev).manual_seed(ci + r)
    x = torch.randn(B, r, r, ci, device=dev, generator=g).bfloat16()
    w = (torch.randn(co, ci, 3, 3, device=dev, generator=g) / (3 * ci ** 0.5))
    wp = K.pack_conv_weight(w, torch.bfloat16, False)
    bias = torch.randn(co, device=dev, generator=g)
    flops = 2.0 * B * r * r * ci * co * 9
    line = f'{kind:5s} {ci:4d}->{co:4d} {r:3d}px:'
    outs = []
    for mode in ('0', '1', '2'):
        os.environ['OOD_CTA2'] = mode
        if kind == 'plain':
            fn = lambda: K.conv3x3(x, wp, co, bias=bias, act=True)
        elif kind == 'stats':
            fn = lambda: K.conv3x3(x, wp, co, bias=bias, stats_eps=1e-5)
        else:
            seed, _ = K.conv3x3(x, wp, co, out_f32=True, tiled=True)
            fn = lambda: K.conv3x3(x, wp, co, bias=bias, acc_in=seed, tiled=True)
        t = timeit(fn)
        outs.append(fn()[0])
        line += f'  CTA2={mode} {t:7.1f} us {flops / t / 1e6:7.1f} TFLOP/s'
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    print(line, flush=True)
