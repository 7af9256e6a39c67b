"""Micro-benchmark of the image byte-format kernels at the bench shape (16 x 1024 x 1024 x 3): 15 bytes per pixel each."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ood_gan_inversion_b200 import kernels as K  # noqa: E402


def timeit(fn, warm=3, rep=10):
    for _ in range(warm):
        fn()
    best = 1e9
    for _ in range(rep):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


b, s = 16, 1024
frames = torch.randint(0, 256, (b, s, s, 3), dtype=torch.uint8, device='cuda')
t = torch.randn(b, 3, s, s, device='cuda')
byt = b * s * s * 15
for name, fn in (('img2tensor_u8', lambda: K.img2tensor_u8(frames)), ('tensor2img_u8', lambda: K.tensor2img_u8(t, lo=-1.0, hi=1.0))):
    ms = timeit(fn)
    print(f'{name} B{b} {s}px: {ms * 1e3:.1f} us  {byt / ms / 1e6:.0f} GB/s  frac {byt / ms / 1e6 / 6534.8:.3f}')
