"""Sweep of the warp_mix launch knobs (OOD_WARP_VPT / OOD_WARP_ITER / OOD_WARP_SEG, read per call) at the config-2 shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ood_gan_inversion_b200 import kernels as K  # noqa: E402
from scripts.samm_bench import timeit  # noqa: E402  (also prints the default-route table first)

b = 16
for c, r in [(512, 32), (512, 64), (256, 128), (128, 256)]:
    gen = torch.randn(b, r, r, c, device='cuda').bfloat16()
    lo = torch.randn(b, 3, max(r // 8, 2), max(r // 8, 2), device='cuda')
    up = torch.nn.functional.interpolate(lo, size=(r, r), mode='bilinear', align_corners=False)
    field = torch.cat([0.08 * torch.tanh(up[:, :2]), torch.sigmoid(up[:, 2:])], 1).contiguous()
    byt = b * (2 * c * r * r * 2 + 3 * r * r * 4)
    ref = None
    res = []
    for vpt in (1, 2, 4):
        for it in (1, 2, 4, 8):
            for seg in (8, 16, 32):
                os.environ.update(OOD_WARP_VPT=str(vpt), OOD_WARP_ITER=str(it), OOD_WARP_SEG=str(seg))
                out = K.warp_mix(gen, field)
                if ref is None:
                    ref = out
                assert torch.equal(out, ref), (vpt, it, seg)
                res.append((timeit(lambda: K.warp_mix(gen, field), warm=2, rep=6), vpt, it, seg))
    for k in ('OOD_WARP_VPT', 'OOD_WARP_ITER', 'OOD_WARP_SEG'):
        os.environ.pop(k)
    dflt = timeit(lambda: K.warp_mix(gen, field))
    res.sort()
    print(f'C{c} R{r}: default {dflt * 1e3:.1f} us ({byt / dflt / 1e6 / 6534.8:.3f}) | best ' +
          ' | '.join(f'vpt{v} it{i} seg{s} {ms * 1e3:.1f} us ({byt / ms / 1e6 / 6534.8:.3f})' for ms, v, i, s in res[:5]))
