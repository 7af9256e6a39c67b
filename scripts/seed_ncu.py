"""Three launches for an ncu capture at one AlignNet level (default C=128, R=256, batch 16): half-K conv (bf16 out, PReLU),
half-K conv -> tile-order fp32 seed, half-K conv seeded with it."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ood_gan_inversion_b200 import kernels as K  # noqa: E402

c, r, B = int(os.environ.get('C', 128)), int(os.environ.get('R', 256)), int(os.environ.get('B', 16))
torch.manual_seed(0)
lo = torch.randn(B, r, r, c, device='cuda').bfloat16()
hi = torch.randn(B, r, r, c, device='cuda').bfloat16()
w = 0.02 * torch.randn(2 * c, 2 * c, 3, 3, device='cuda')
wl = K.pack_conv_weight(w[:, :c].contiguous(), torch.bfloat16, False)
wh = K.pack_conv_weight(w[:, c:].contiguous(), torch.bfloat16, False)
slope = torch.full((2 * c,), 0.25, device='cuda')
for _ in range(2):
    K.conv3x3(lo, wl, 2 * c, prelu=slope)
    seed, _ = K.conv3x3(hi, wh, 2 * c, out_f32=True, tiled=True)
    K.conv3x3(lo, wl, 2 * c, prelu=slope, acc_in=seed, tiled=True)
torch.cuda.synchronize()
