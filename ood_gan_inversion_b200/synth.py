"""Synthetic (random-init) weights and inputs for benchmarks and smoke runs: there are no checkpoints offline.

Follows SURVEY.md section 8(d): reference init distributions, then the zero-init parameters that would make the path
degenerate (NoiseInjection.weight = 0 => 0/0 in the reference's callback, finding 7) are made non-zero.
"""
import torch
from torch.nn import functional as F


@torch.no_grad()
def synthetic_init(net, seed=0, rgb_gain=0.1):
    g = torch.Generator().manual_seed(seed)
    for name, p in net.named_parameters():
        if name.endswith('noise.weight'):
            p.fill_(0.1)
        elif name.endswith('activate.bias') or (name.endswith('.bias') and 'to_rgb' in name and p.dim() == 4):
            p.copy_(0.1 * torch.randn(p.shape, generator=g))
        elif 'to_rgb' in name and name.endswith('conv.weight'):
            p.mul_(rgb_gain)
        elif name == 'avg_latent':
            p.copy_(0.1 * torch.randn(p.shape, generator=g))
    return net


def synthetic_faces(batch, size=1024, seed=2, device='cpu', pin=False):
    """Smooth, deterministic [-1,1] images (bicubic-upsampled 64x64 Gaussian field), fp32 NCHW."""
    x = torch.randn(batch, 3, 64, 64, generator=torch.Generator().manual_seed(seed))
    x = F.interpolate(x, (size, size), mode='bicubic', align_corners=False).clamp_(-1, 1)
    if pin:
        x = x.pin_memory()
    return x.to(device)
