"""upfirdn2d on fp32 NCHW planes at the large config-5 sizes: blur / up2 / down2, GB/s against the measured HBM peak."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ood_gan_inversion_b200 import kernels as K  # noqa: E402


def timeit(fn, warm=3, rep=8):
    for _ in range(warm):
        fn()
    best = 1e9
    for _ in range(rep):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


k4 = torch.tensor([1., 3., 3., 1.], device='cuda')
k2d = torch.outer(k4, k4) / 64
for res, planes in [(256, 2048), (512, 1024), (1024, 256), (1024, 1024), (2048, 256)]:
    for name, shape, args in [('blur', (planes, res + 1, res + 1), (1, 1, 1, 1, 1, 1, 1, 1)),
                              ('up2', (planes, res // 2, res // 2), (2, 2, 1, 1, 2, 1, 2, 1)),
                              ('down2', (planes, res, res), (1, 1, 2, 2, 1, 1, 1, 1))]:
        x = torch.randn(1, *shape, device='cuda')
        y = K.upfirdn2d_nchw(x, k2d, *args)
        ms = timeit(lambda: K.upfirdn2d_nchw(x, k2d, *args))
        byt = (x.numel() + y.numel()) * 4
        print(f'{name:6s} fp32 res {res:5d} planes {planes:5d}: {ms * 1e3:8.1f} us  {byt / ms / 1e6:7.0f} GB/s  frac {byt / ms / 1e6 / 6534.8:.3f}')
