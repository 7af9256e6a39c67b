"""A few launches of the feats_conv-shaped 1x1 convolution (64 -> 128 at 256 px, batch 16, f16 in / bf16 out, bias) for ncu, and its timing."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ood_gan_inversion_b200 import kernels as K  # noqa: E402
b, r, ci, co = 16, 256, 64, 128
x = torch.randn(b, r, r, ci, device='cuda').half()
w = K.pack_conv1x1_weight(torch.randn(co, ci, 1, 1, device='cuda') * 0.1, torch.float16, False)
bias = torch.randn(co, device='cuda')
fn = lambda: K.conv3x3(x, w, co, transposed=4, bias=bias, out_dtype=torch.bfloat16)
for _ in range(4):
    fn()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
byt = b * r * r * (ci + co) * 2
print(f'conv1x1 {ci}->{co} {r}px: median {ts[5]:.1f} us  {byt / ts[5] / 1e3:.0f} GB/s ({byt / ts[5] / 1e3 / 6534.8:.2f} of HBM)')
