"""A few launches of ood_se_apply at the encoder's stage-3 shape (256 channels, 32 px, batch 16, f16, fp32 shortcut) for ncu."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ood_gan_inversion_b200 import kernels as K  # noqa: E402
B, c, r = 16, 256, 32
g = torch.Generator(device='cuda').manual_seed(0)
u = torch.randn(B, r, r, c, device='cuda', generator=g).half()
wp = K.pack_conv_weight(torch.randn(c, c, 3, 3, device='cuda', generator=g) / 48, torch.float16, False)
bias = torch.randn(c, device='cuda', generator=g)
w1 = torch.randn(c // 16, c, device='cuda', generator=g) / 16
w2 = torch.randn(c, c // 16, device='cuda', generator=g) / 4
sc = torch.randn(B, r, r, c, device='cuda', generator=g)
bn_g, bn_h = torch.rand(c, device='cuda', generator=g) + 0.5, torch.randn(c, device='cuda', generator=g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for _ in range(5):
    v, _, sums = K.conv3x3(u, wp, c, bias=bias, tile_sums=True)
    flush.zero_()
    K.se_apply(v, sums, w1, w2, sc, 1, bn_g, bn_h)
torch.cuda.synchronize()
