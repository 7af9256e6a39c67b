import sys, torch
sys.path.insert(0, '.')
from ood_gan_inversion_b200 import kernels as K
k4 = torch.tensor([1., 3., 3., 1.], device='cuda'); k2d = torch.outer(k4, k4) / 64
for dt in (torch.float32, torch.bfloat16):
    x = torch.randn(4, 64, 1025, 1025, device='cuda').to(dt)
    for _ in range(3):
        K.upfirdn2d_nchw(x, k2d * 4, 1, 1, 1, 1, 1, 1, 1, 1)
    x2 = torch.randn(4, 64, 512, 512, device='cuda').to(dt)
    for _ in range(3):
        K.upfirdn2d_nchw(x2, k2d * 4, 2, 2, 1, 1, 2, 1, 2, 1)
torch.cuda.synchronize()
