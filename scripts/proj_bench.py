"""AlignNet head micro-benchmark: per-sample grouped 1x1 projection (2C -> 32, fp32 out) + tap_sum at the four SAMM levels."""
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
for c2, r in [(1024, 32), (1024, 64), (512, 128), (256, 256)]:
    x = torch.randn(b, r, r, c2, device='cuda').bfloat16()
    w = (torch.randn(b, 32, c2, device='cuda') * 0.05).bfloat16()
    bias = torch.randn(b, 32, device='cuda')
    for groups in (b, 1):
        ww = w if groups == b else w[:1].contiguous()
        bb = bias if groups == b else bias[0].contiguous()
        ms = timeit(lambda: K.conv3x3(x, ww, 32, transposed=4, out_f32=True, groups=groups, bias=bb))
        byt = b * r * r * (c2 * 2 + 32 * 4)
        print(f'proj 2C={c2} R{r} groups={groups}: {ms * 1e3:.1f} us  {byt / ms / 1e6:.0f} GB/s  frac {byt / ms / 1e6 / 6534.8:.3f}')
    p = torch.randn(b, r, r, 32, device='cuda')
    ms = timeit(lambda: K.tap_sum(p))
    print(f'tap_sum R{r}: {ms * 1e3:.1f} us')
