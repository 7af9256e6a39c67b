"""Oracle (test infrastructure): Spatial Alignment & Masking Module, functional.

Restates src/ops/SAMM/helpers.py (new_PRM :62-77, AlignNet :85-109, SPM_Warp :111-179) and
bottleneck_IR / BN of src/ops/e4e/encoders/helpers.py:93-99,426-448 of the reference.
"""
import torch
import torch.nn.functional as F

from .ops import upfirdn2d


def instance_norm(x, weight=None, bias=None, eps=1e-5):
    """nn.InstanceNorm2d (biased variance, no running stats).  helpers.py:88, e4e helpers.py:94-95."""
    mu = x.mean(dim=(2, 3), keepdim=True)
    var = x.var(dim=(2, 3), unbiased=False, keepdim=True)
    y = (x - mu) * torch.rsqrt(var + eps)
    if weight is not None:
        y = y * weight.reshape(1, -1, 1, 1) + bias.reshape(1, -1, 1, 1)
    return y


def bottleneck_ir_in(sd, p, x):
    """bottleneck_IR with affine InstanceNorm and bias-free convs.  e4e helpers.py:426-448."""
    if f'{p}shortcut_layer.0.weight' in sd:
        sc = F.conv2d(x, sd[f'{p}shortcut_layer.0.weight'])
        sc = instance_norm(sc, sd[f'{p}shortcut_layer.1.weight'], sd[f'{p}shortcut_layer.1.bias'])
    else:
        sc = x  # MaxPool2d(1, 1) is the identity
    r = instance_norm(x, sd[f'{p}res_layer.0.weight'], sd[f'{p}res_layer.0.bias'])
    r = F.conv2d(r, sd[f'{p}res_layer.1.weight'], padding=1)
    r = F.prelu(r, sd[f'{p}res_layer.2.weight'])
    r = F.conv2d(r, sd[f'{p}res_layer.3.weight'], padding=1)
    r = instance_norm(r, sd[f'{p}res_layer.4.weight'], sd[f'{p}res_layer.4.bias'])
    return r + sc


def align_net(sd, p, cur, enc, scale):
    """[tanh*scale, tanh*scale, sigmoid] of body(cat[IN(cur)-IN(enc), IN(enc)]).  helpers.py:96-109."""
    a, e = instance_norm(cur), instance_norm(enc)
    z = torch.cat([a - e, e], dim=1)
    z = bottleneck_ir_in(sd, f'{p}body.0.', z)
    z = bottleneck_ir_in(sd, f'{p}body.1.', z)
    return torch.cat([torch.tanh(z[:, 0:1]) * scale, torch.tanh(z[:, 1:2]) * scale,
                      torch.sigmoid(z[:, 2:])], dim=1)


def prm(x, y):
    """new_PRM: y*up(x) + up(x)*(1-up(x)), bicubic align_corners=True when sizes differ.  :62-77."""
    if x.shape[-2:] != y.shape[-2:]:
        x = F.interpolate(x, size=y.shape[-2:], mode='bicubic', align_corners=True)
    return y * x + x * (1 - x)


def warp_mix(gen, field):
    """grid = linspace(-1,1) base + (dx,dy); bilinear/zeros/align_corners=False sample; alpha mix.
    helpers.py:168-177."""
    b, _, h, w = gen.shape
    ly = torch.linspace(-1, 1, h, device=gen.device)
    lx = torch.linspace(-1, 1, w, device=gen.device)
    gx = lx.reshape(1, 1, w) + field[:, 0]
    gy = ly.reshape(1, h, 1) + field[:, 1]
    grid = torch.stack([gx, gy], dim=-1)
    warped = F.grid_sample(gen, grid, mode='bilinear', padding_mode='zeros', align_corners=False)
    alpha = field[:, 2:]
    return warped * alpha + gen * (1 - alpha)


def spm_warp(sd, p, enc, gen, coarse=None, scale=0.08, cycles=2):
    """Iterative alignment.  helpers.py:149-179.  Returns (aligned features, field [B,3,R,R])."""
    cur, acc = gen, None
    for k in range(cycles):
        f = upfirdn2d(align_net(sd, f'{p}body.', cur, enc, scale), sd[f'{p}blur.kernel'], pad=(2, 1))
        if acc is None:
            acc = f
        else:
            acc = torch.cat([torch.clip(acc[:, 0:1] + f[:, 0:1], -scale, scale),
                             torch.clip(acc[:, 1:2] + f[:, 1:2], -scale, scale),
                             torch.clip(prm(acc[:, 2:], f[:, 2:]), 0.0, 1.0)], dim=1)
        if k == cycles - 1 and coarse is not None:
            acc = torch.cat([acc[:, 0:2], torch.clip(prm(coarse[:, 2:], acc[:, 2:]), 0.0, 1.0)], dim=1)
        cur = warp_mix(gen, acc)
    return cur, acc


def compose_masks(fields, size):
    """Bilinear (align_corners=False) upsample of each level's alpha, PRM-style composition in
    ascending level order, clip.  src/archs/OOD_faceGAN_e4e_arch.py:315-339."""
    a = None
    for f in fields:
        u = F.interpolate(f[:, 2:], size=(size, size), mode='bilinear')
        a = u if a is None else u * a + a * (1 - a)
    return torch.clip(a, 0.0, 1.0)


def blend(alpha, x, gen):
    """e4e_arch.py:341-347."""
    return alpha * x + gen * (1 - alpha)
