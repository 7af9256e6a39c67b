"""ToRGB micro-benchmark at the generator's sizes (bf16 NHWC activations -> fp32 NCHW image with the upsampled skip)."""
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


b = 16
for c, r in [(32, 1024), (64, 512), (128, 256), (256, 256), (512, 64)]:
    y = torch.randn(b, r, r, c, device='cuda').bfloat16()
    wrgb = torch.randn(b, 3, c, device='cuda') * 0.1
    bias = torch.randn(3, device='cuda')
    skip = torch.randn(b, 3, r // 2, r // 2, device='cuda')
    taps = K.fir_taps(gain=2.0)
    ms = timeit(lambda: K.torgb(y, wrgb, bias, skip, taps))
    byt = b * r * r * (c * 2 + 3 * 4) + b * 3 * (r // 2) ** 2 * 4
    print(f'torgb bf16 C{c} R{r}: {ms * 1e3:.1f} us  {byt / ms / 1e6:.0f} GB/s  frac {byt / ms / 1e6 / 6534.8:.3f}')
