"""One launch of the fused-phase transposed convolution at the 1024 px layer (64 -> 32, batch 16) for an ncu capture."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ood_gan_inversion_b200 import kernels as K  # noqa: E402

ci, co, r, b = 64, 32, 512, 16
x = torch.randn(b, r, r, ci, device='cuda').bfloat16()
w = torch.randn(co, ci, 3, 3, device='cuda') * 0.1
wf = K.pack_convt_fused(K.pack_conv_weight(w, torch.bfloat16, False))
for _ in range(3):
    K.conv3x3(x, wf, co, transposed=5)
torch.cuda.synchronize()
