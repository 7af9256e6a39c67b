"""Per-stage timing of the encoder's squeeze-excite tail (batch 16, f16 storage, fp32 residual stream): the cluster kernel (ood_se_tail)
against the convolution-epilogue sums + ood_se_apply, and the cost of the sums in the convolution itself."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ood_gan_inversion_b200 import kernels as K  # noqa: E402

dev = 'cuda'
B = 16
HBM = 6534.8


def timeit(fn, n=30):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[0], ts[len(ts) // 2]


for c, r in ((64, 128), (128, 64), (256, 32), (512, 16)):
    g = torch.Generator(device=dev).manual_seed(c)
    u = torch.randn(B, r, r, c, device=dev, generator=g).half()
    w = (torch.randn(c, c, 3, 3, device=dev, generator=g) / (3 * c ** 0.5)).half()
    wp = K.pack_conv_weight(w.float(), torch.float16, False)
    bias = torch.randn(c, device=dev, generator=g)
    w1 = torch.randn(c // 16, c, device=dev, generator=g) / c ** 0.5
    w2 = torch.randn(c, c // 16, device=dev, generator=g) / 4
    sc = torch.randn(B, r, r, c, device=dev, generator=g)
    bn_g, bn_h = torch.rand(c, device=dev, generator=g) + 0.5, torch.randn(c, device=dev, generator=g)
    v, _ = K.conv3x3(u, wp, c, bias=bias)
    nbytes = v.numel() * (2 + 4 + 4 + 2)
    t_tail = timeit(lambda: K.se_tail(v, w1, w2, sc, 1, bn_g, bn_h))
    line = f'C={c:4d} {r:4d}px: se_tail {t_tail[0]:7.1f} / {t_tail[1]:7.1f} us ({nbytes / t_tail[1] / 1e3 / HBM:.2f} of HBM)'
    t_conv = timeit(lambda: K.conv3x3(u, wp, c, bias=bias))
    line += f' | conv {t_conv[1]:7.1f} us'
    if c % 128 == 0 and K.conv3x3_stats_ok(u, c, 0):
        v2, _, sums = K.conv3x3(u, wp, c, bias=bias, tile_sums=True)
        assert torch.equal(v, v2)
        t_convs = timeit(lambda: K.conv3x3(u, wp, c, bias=bias, tile_sums=True))
        t_app = timeit(lambda: K.se_apply(v, sums, w1, w2, sc, 1, bn_g, bn_h))
        a, b_ = K.se_tail(v, w1, w2, sc, 1, bn_g, bn_h), K.se_apply(v, sums, w1, w2, sc, 1, bn_g, bn_h)
        err = float((a[0] - b_[0]).abs().max())
        line += f' conv+sums {t_convs[1]:7.1f} us | se_apply {t_app[0]:7.1f} / {t_app[1]:7.1f} us ({nbytes / t_app[1] / 1e3 / HBM:.2f} of HBM)  max diff {err:.2e}'
    print(line, flush=True)
