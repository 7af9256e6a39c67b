"""Transposed (stride-2) modulated-conv micro-benchmark at the generator's sizes (bf16 out, raw accumulators)."""
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
for ci, co, r in [(64, 32, 512), (128, 64, 256), (256, 128, 128), (512, 256, 64), (512, 512, 32)]:
    x = torch.randn(b, r, r, ci, device='cuda').bfloat16()
    w = torch.randn(co, ci, 3, 3, device='cuda') * 0.1
    wp = K.pack_conv_weight(w, torch.bfloat16, False)
    ms = timeit(lambda: K.conv3x3(x, wp, co, transposed=True))
    byt = b * (r * r * ci + (2 * r + 1) ** 2 * co) * 2
    fl = 2.0 * b * co * ci * 9 * r * r
    print(f'convT {ci}->{co} {r}->{2 * r + 1}: {ms * 1e3:.1f} us  {byt / ms / 1e6:.0f} GB/s ({byt / ms / 1e6 / 6534.8:.3f} of HBM)  {fl / ms / 1e9:.0f} TFLOP/s')
    wf = K.pack_convt_fused(wp)
    ms = timeit(lambda: K.conv3x3(x, wf, co, transposed=5))
    print(f'   fused phases (form 5):        {ms * 1e3:.1f} us  {byt / ms / 1e6:.0f} GB/s ({byt / ms / 1e6 / 6534.8:.3f} of HBM)  {fl / ms / 1e9:.0f} TFLOP/s')
