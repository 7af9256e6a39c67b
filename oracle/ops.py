"""Oracle (test infrastructure): FIR resampling and bias+leaky-ReLU.

Restates src/ops/op/upfirdn2d.py:160-193 (upfirdn2d_native) and
src/ops/op/fused_act.py:92-96 (native fused_leaky_relu) of the reference.
"""
import math

import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2.0)


def fir_kernel(taps, gain=1.0):
    """Normalised 2-D FIR from 1-D (outer product) or 2-D taps.  src/ops/StyleGAN/model.py:19-27."""
    k = torch.as_tensor(taps, dtype=torch.float32)
    if k.dim() == 1:
        k = torch.outer(k, k)
    return k / k.sum() * gain


def upfirdn2d(x, k, up=1, down=1, pad=(0, 0)):
    """Python-API form: the same (pad0, pad1) on both axes.  src/ops/op/upfirdn2d.py:149-157."""
    return upfirdn2d_xy(x, k, up, up, down, down, pad[0], pad[1], pad[0], pad[1])


def upfirdn2d_xy(x, k, up_x, up_y, down_x, down_y, px0, px1, py0, py1):
    """out = decimate(correlate(pad(zero_insert(x)), flip(k))).  src/ops/op/upfirdn2d.py:160-193.

    x: [B,C,H,W]; zeros are inserted AFTER each sample; negative pads crop.
    """
    b, c, h, w = x.shape
    kh, kw = k.shape
    planes = x.reshape(b * c, 1, h, w)
    if up_x > 1 or up_y > 1:
        stuffed = planes.new_zeros(b * c, 1, h * up_y, w * up_x)
        stuffed[:, :, ::up_y, ::up_x] = planes
        planes = stuffed
    planes = F.pad(planes, [px0, px1, py0, py1])  # negative values crop
    taps = torch.flip(k, [0, 1]).to(planes.dtype).reshape(1, 1, kh, kw)
    full = F.conv2d(planes, taps)
    full = full[:, :, ::down_y, ::down_x]
    return full.reshape(b, c, full.shape[2], full.shape[3])


def upfirdn2d_out_size(n, up, down, p0, p1, kn):
    """src/ops/op/upfirdn2d.py:107-108."""
    return (n * up + p0 + p1 - kn) // down + 1


def fused_leaky_relu(x, bias, negative_slope=0.2, scale=SQRT2):
    """scale * lrelu(x + bias[channel]).  src/ops/op/fused_act.py:96."""
    if bias is not None:
        x = x + bias.reshape((1, -1) + (1,) * (x.dim() - 2))
    return F.leaky_relu(x, negative_slope) * scale


def fused_leaky_relu_backward(grad_out, out, negative_slope=0.2, scale=SQRT2):
    """Gradient gated on the sign of the saved OUTPUT.  src/ops/op/fused_act.py:27-45 and
    src/ops/op/fused_bias_act_kernel.cu:36-47 (act=3, grad=1)."""
    gate = torch.where(out > 0, torch.ones_like(out), torch.full_like(out, negative_slope))
    gx = grad_out * gate * scale
    dims = [0] + list(range(2, gx.dim()))
    return gx, gx.sum(dims)
