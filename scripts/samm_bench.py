"""Micro-benchmark of the SAMM bandwidth kernels (warp + alpha mix, mask compose + blend) at the config-2 shapes."""
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
for c, r in [(512, 32), (512, 64), (256, 128), (128, 256)]:
    gen = torch.randn(b, r, r, c, device='cuda').bfloat16()
    # smooth flow (the pipeline's fields are FIR-blurred): low-resolution noise, bilinearly upsampled, +-0.08
    lo = torch.randn(b, 3, max(r // 8, 2), max(r // 8, 2), device='cuda')
    up = torch.nn.functional.interpolate(lo, size=(r, r), mode='bilinear', align_corners=False)
    field = torch.cat([0.08 * torch.tanh(up[:, :2]), torch.sigmoid(up[:, 2:])], 1).contiguous()
    ms = timeit(lambda: K.warp_mix(gen, field))
    byt = b * (2 * c * r * r * 2 + 3 * r * r * 4)
    print(f'warp_mix bf16 C{c} R{r}: {ms * 1e3:.1f} us  {byt / ms / 1e6:.0f} GB/s  frac {byt / ms / 1e6 / 6534.8:.3f}')
size = 1024
fields = [torch.rand(b, 3, s, s, device='cuda') for s in (32, 64, 128, 256)]
x, g = torch.randn(b, 3, size, size, device='cuda'), torch.randn(b, 3, size, size, device='cuda')
ms = timeit(lambda: K.mask_blend(fields, x, g))
byt = b * (3 * 3 * size * size * 4 + size * size * 4)
print(f'mask_blend B{b}: {ms * 1e3:.1f} us  {byt / ms / 1e6:.0f} GB/s  frac {byt / ms / 1e6 / 6534.8:.3f}')

# ---- backward kernels (gradient path): time, algorithmic bytes, and run-to-run repeatability of the atomic accumulations
print('backward kernels:')
for c, r in [(512, 64), (128, 256)]:
    gen = torch.randn(b, r, r, c, device='cuda').bfloat16()
    lo = torch.randn(b, 3, max(r // 8, 2), max(r // 8, 2), device='cuda')
    up = torch.nn.functional.interpolate(lo, size=(r, r), mode='bilinear', align_corners=False)
    field = torch.cat([0.08 * torch.tanh(up[:, :2]), torch.sigmoid(up[:, 2:])], 1).contiguous()
    gout = torch.randn_like(gen)
    ms = timeit(lambda: K.warp_mix_bwd(gen, field, gout))
    byt = b * (2 * c * r * r * 2 + c * r * r * 4 + 2 * 3 * r * r * 4)          # gen + gout read (bf16), ggen written (fp32), field read, gfield written
    g1, f1 = K.warp_mix_bwd(gen, field, gout)
    g2, f2 = K.warp_mix_bwd(gen, field, gout)
    print(f'warp_mix_bwd bf16 C{c} R{r}: {ms * 1e3:.1f} us  {byt / ms / 1e6:.0f} GB/s  frac {byt / ms / 1e6 / 6534.8:.3f};  run-to-run max |diff| ggen '
          f'{float((g1 - g2).abs().max()):.3g} (of {float(g1.abs().max()):.3g}), gfield {float((f1 - f2).abs().max()):.3g} (of {float(f1.abs().max()):.3g})')
gout = torch.randn(b, 3, size, size, device='cuda')
ms = timeit(lambda: K.mask_blend_bwd(fields, x, g, gout, want_gx=False, want_ggen=True))
byt = b * (4 * 3 * size * size * 4)
r1 = K.mask_blend_bwd(fields, x, g, gout, want_gx=False, want_ggen=True)
r2 = K.mask_blend_bwd(fields, x, g, gout, want_gx=False, want_ggen=True)
d = max(float((a - c_).abs().max()) for a, c_ in zip(r1[2], r2[2]))
print(f'mask_blend_bwd B{b}: {ms * 1e3:.1f} us  {byt / ms / 1e6:.0f} GB/s  frac {byt / ms / 1e6 / 6534.8:.3f};  run-to-run max |diff| of the field gradients {d:.3g} '
      f'(of {max(float(a.abs().max()) for a in r1[2]):.3g})')
