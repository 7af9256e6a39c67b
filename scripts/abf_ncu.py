"""A few launches of ood_act_bwd_fused at config 4's top level (batch 32, 1024 px, 32 channels, bf16, ToRGB gradient) for ncu, and its timing."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ood_gan_inversion_b200 import kernels as K  # noqa: E402
b, r, c = 32, 1024, 32
g = torch.Generator(device='cuda').manual_seed(0)
y = torch.randn(b, r, r, c, device='cuda', generator=g).bfloat16()
g_in = torch.randn(b, r, r, c, device='cuda', generator=g).bfloat16()
g_rgb = torch.randn(b, 3, r, r, device='cuda', generator=g)
wrgb = torch.randn(b, 3, c, device='cuda', generator=g)
d, s = torch.rand(b, c, device='cuda', generator=g) + 0.5, torch.rand(b, c, device='cuda', generator=g) + 0.5
bias, noise, nw = torch.randn(c, device='cuda', generator=g), torch.randn(b, 1, r, r, device='cuda', generator=g), torch.tensor([0.1], device='cuda')
fn = lambda: K.act_bwd_fused(g_in, s, (g_rgb, wrgb), y, d, bias, noise, nw)
for _ in range(3):
    fn()
torch.cuda.synchronize()
ts = []
for _ in range(8):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
byt = b * r * r * (3 * c * 2 + 16)
print(f'act_bwd_fused {c} ch {r} px batch {b}: median {ts[4]:.1f} us  {byt / ts[4] / 1e3:.0f} GB/s ({byt / ts[4] / 1e3 / 6534.8:.2f} of HBM)')
