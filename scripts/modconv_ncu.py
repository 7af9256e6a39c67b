"""A few launches of generator- / AlignNet-shaped convolutions for ncu (alignnet: 1024 -> 1024, 64 px, fused statistics, CTA pairs): the 512 -> 512 modulated conv at 64 px, plain (form 0, full StyledConv
epilogue) and transposed (form 1), batch 16; and the AlignNet weight-gradient GEMM (1024 -> 1024, 64 px, batch 4)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ood_gan_inversion_b200 import kernels as K  # noqa: E402
which = sys.argv[1] if len(sys.argv) > 1 else 'plain'
b, r, c = 16, 64, 512
x = torch.randn(b, r, r, c, device='cuda').bfloat16()
w = torch.randn(c, c, 3, 3, device='cuda') * 0.05
wp = K.pack_conv_weight(w, torch.bfloat16, False)
d, bias, sn = torch.rand(b, c, device='cuda') + 0.5, torch.randn(c, device='cuda'), torch.rand(b, c, device='cuda') + 0.5
noise, nw = torch.randn(b, 1, r, r, device='cuda'), torch.tensor([0.1], device='cuda')
if which == 'alignnet':      # the dominant kernel of the step: AlignNet 1024 -> 1024 at 64 px with fused statistics, CTA pairs (conv_tc_kernel<256,64,STATS,CTA2>)
    c2 = 1024
    xa = torch.randn(b, r, r, c2, device='cuda').bfloat16()
    wa = K.pack_conv_weight(torch.randn(c2, c2, 3, 3, device='cuda') * 0.02, torch.bfloat16, False)
for _ in range(4):
    if which == 'alignnet':
        K.conv3x3(xa, wa, c2, stats_eps=1e-5)
    elif which == 'plain':
        K.conv3x3(x, wp, c, d=d, noise=noise, noise_w=nw, bias=bias, s_next=sn, act=True, want_y=False, want_ys=True)
    elif which == 'transposed':
        K.conv3x3(x, wp, c, transposed=True)
    else:
        g = torch.randn(4, 64, 64, 1024, device='cuda').bfloat16()
        xx = torch.randn(4, 64, 64, 1024, device='cuda').bfloat16()
        K.conv_wgrad(g, xx, 9)
torch.cuda.synchronize()
