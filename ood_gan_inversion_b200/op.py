"""Drop-in for the reference's `src.ops.op` package: upfirdn2d, fused_leaky_relu, FusedLeakyReLU.

Same names, argument meaning and autograd behaviour (first and second order) as
src/ops/op/upfirdn2d.py:23-157 and src/ops/op/fused_act.py:25-96 of the reference, but every call runs the
sm_100a kernels behind the C ABI (ood_upfirdn2d / ood_fused_bias_act).  The reference's `device='cpu'` switch
(which silently selects its native PyTorch branch, SURVEY finding 1) is accepted and ignored: CUDA tensors run
the kernels, anything else raises.
"""
import math

import torch
from torch import nn
from torch.autograd import Function

from . import kernels as K


class _UpFirDn2dBackward(Function):
    """src/ops/op/upfirdn2d.py:23-89: gradient = the same op with the flipped FIR, up/down swapped, g_pad."""

    @staticmethod
    def forward(ctx, grad_output, kernel, grad_kernel, up, down, pad, g_pad, in_size, out_size):
        gx = K.upfirdn2d_nchw(grad_output.reshape(in_size[0], in_size[1], out_size[0], out_size[1]), grad_kernel,
                              down[0], down[1], up[0], up[1], *g_pad)
        # the adjoint may be larger than the input when (in*up + pads - k) is not a multiple of down
        gx = gx[:, :, :in_size[2], :in_size[3]] if gx.shape[2:] != tuple(in_size[2:]) else gx
        ctx.save_for_backward(kernel)
        ctx.cfg = (up, down, pad, in_size, out_size)
        return gx.contiguous()

    @staticmethod
    def backward(ctx, gradgrad_input):
        kernel, = ctx.saved_tensors
        up, down, pad, in_size, out_size = ctx.cfg
        ggo = K.upfirdn2d_nchw(gradgrad_input.contiguous(), kernel, up[0], up[1], down[0], down[1], *pad)
        return ggo, None, None, None, None, None, None, None, None


class _UpFirDn2d(Function):
    """src/ops/op/upfirdn2d.py:92-146."""

    @staticmethod
    def forward(ctx, input, kernel, up, down, pad):
        up_x, up_y = up
        down_x, down_y = down
        px0, px1, py0, py1 = pad
        kh, kw = kernel.shape
        _, _, in_h, in_w = input.shape
        out = K.upfirdn2d_nchw(input, kernel, up_x, up_y, down_x, down_y, px0, px1, py0, py1)
        out_h, out_w = out.shape[2:]
        ctx.save_for_backward(kernel, torch.flip(kernel, [0, 1]))
        ctx.cfg = (up, down, pad, tuple(input.shape), (out_h, out_w))
        # upfirdn2d.py:115-118
        ctx.g_pad = (kw - px0 - 1, in_w * up_x - out_w * down_x + px0 - up_x + 1,
                     kh - py0 - 1, in_h * up_y - out_h * down_y + py0 - up_y + 1)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        kernel, grad_kernel = ctx.saved_tensors
        up, down, pad, in_size, out_size = ctx.cfg
        gx = _UpFirDn2dBackward.apply(grad_output.contiguous(), kernel, grad_kernel, up, down, pad, ctx.g_pad, in_size, out_size)
        return gx, None, None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0), device='cpu'):
    """src/ops/op/upfirdn2d.py:149-157 (same signature; `device` ignored, see module docstring)."""
    return _UpFirDn2d.apply(input, kernel, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]))


class _FusedLeakyReLUBackward(Function):
    """src/ops/op/fused_act.py:25-54."""

    @staticmethod
    def forward(ctx, grad_output, out, negative_slope, scale):
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        grad_input = K.fused_bias_act(grad_output.contiguous(), None, out, 1, negative_slope, scale)
        if grad_input.dim() > 1:
            grad_bias = K.bias_grad(grad_input)
        else:
            grad_bias = grad_input.float()
        return grad_input, grad_bias

    @staticmethod
    def backward(ctx, gradgrad_input, gradgrad_bias):
        out, = ctx.saved_tensors
        # d/d(grad_output) of grad_input, plus the bias path (fused_act.py:47-54: the gate is applied to gg_in + gg_bias)
        gg = gradgrad_input
        if gradgrad_bias is not None:
            gg = gg + gradgrad_bias.to(gg.dtype).reshape((1, -1) + (1,) * (gg.dim() - 2))
        ggo = K.fused_bias_act(gg.contiguous(), None, out, 1, ctx.negative_slope, ctx.scale)
        return ggo, None, None, None


class _FusedLeakyReLU(Function):
    """src/ops/op/fused_act.py:57-76."""

    @staticmethod
    def forward(ctx, input, bias, negative_slope, scale):
        out = K.fused_bias_act(input, bias, None, 0, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        ctx.bias_dtype = bias.dtype
        return out

    @staticmethod
    def backward(ctx, grad_output):
        out, = ctx.saved_tensors
        gi, gb = _FusedLeakyReLUBackward.apply(grad_output, out, ctx.negative_slope, ctx.scale)
        return gi, gb.to(ctx.bias_dtype), None, None


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5, device='cpu'):
    """scale * leaky_relu(input + bias[channel]).  src/ops/op/fused_act.py:92-96."""
    return _FusedLeakyReLU.apply(input, bias, negative_slope, scale)


class FusedLeakyReLU(nn.Module):
    """src/ops/op/fused_act.py:79-89: same constructor, `.bias` parameter and state-dict key."""

    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5, device='cpu'):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope = negative_slope
        self.scale = scale
        self.device = device

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale, self.device)


SQRT2 = math.sqrt(2.0)
