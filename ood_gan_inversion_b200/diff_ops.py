"""Differentiable building blocks over this library's kernels.

The inference path fuses everything it can (epilogues that carry the next layer's style, statistics from the convolution,
folded InstanceNorms).  The gradient path needs the intermediate values, so it is composed from small autograd Functions
whose forward AND backward are kernels of libood_b200 (there is no torch fallback for a 2C-channel tensor):

    InstNorm      InstanceNorm2d (affine or not) on NHWC;   backward = two reductions + one combine pass (ood_nhwc_affine2)
    Conv          shared-weight 3x3 / 1x1 convolution;        backward = the same implicit-GEMM kernel on the transposed pack
    PReLU         ood_prelu / its backward
    Add / Sub / Cat   ood_nhwc_affine2 on channel slices
    Head27 / Shortcut3   the AlignNet's 2C -> 3 3x3 head (projection + tap_sum; backward tap_gather + 1x1) and its 1x1 shortcut

Reference: torch.autograd through src/ops/SAMM/helpers.py:85-109 (AlignNet), e4e/encoders/helpers.py:426-448 (bottleneck_IR).
Weight gradients (the training step of src/models/OOD_faceGAN_model.py:663-789): InstNorm / Shortcut3 / bias
gradients fall out of the same reductions; Conv / Head27 use the tcgen05 pixel-contraction kernel (ood_conv_wgrad: both operands
MN-major straight from NHWC, deterministic K slices); PReLU slopes = one extra reduction.  They are computed only for parameters that
require grad, so the latent-only inversion path pays nothing for them.
"""
import torch
from torch.autograd import Function

from . import kernels as K
from . import stylegan as sg


def _needs(t):
    return t is not None and isinstance(t, torch.Tensor) and t.requires_grad


class _InstNorm(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        x = x.contiguous()
        st = K.in_stats(x, None, eps)
        ctx.save_for_backward(x, st, weight)
        ctx.has_bias = bias is not None
        return K.in_apply(x, st, None if weight is None else weight.detach().float().contiguous(),
                          None if bias is None else bias.detach().float().contiguous())

    @staticmethod
    def backward(ctx, g):
        x, st, weight = ctx.saved_tensors
        g = g.contiguous().to(x.dtype)
        b, h, w, c = x.shape
        n = float(h * w)
        mu, r = st[..., 0], st[..., 1]
        m1 = K.in_stats(g)[..., 0]                               # mean(g)
        sgx = K.dot_reduce(g, x)                                  # sum(g * x)
        m2 = r * (sgx / n - mu * m1)                              # mean(g * xhat)
        gam = weight.detach().float().reshape(1, -1) if weight is not None else 1.0
        a = (r * gam).contiguous()
        gx = K.affine2(g, x, a, (-a * r * m2).contiguous(), (a * (r * m2 * mu - m1)).contiguous())
        gw = (m2 * n).sum(0) if (weight is not None and ctx.needs_input_grad[1]) else None
        gb = (m1 * n).sum(0) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return gx, gw, gb, None


def inst_norm(x, weight=None, bias=None, eps=1e-5):
    """NHWC InstanceNorm2d: (x - mean_hw) * rsqrt(var_hw + eps) * weight + bias."""
    return _InstNorm.apply(x, weight, bias, eps)


_pack_cache = {}


def _packs(weight, kind):
    """(forward pack, data-gradient pack) of a Conv2d weight [Co,Ci,k,k] for the active precision, cached on (id, version)."""
    key = (id(weight), weight._version, weight.data_ptr(), sg.get_precision(), kind)
    hit = _pack_cache.get(key)
    if hit is None:
        if len(_pack_cache) > 64:
            _pack_cache.clear()
        with torch.no_grad():
            dt, cim = sg._act_dtype(), sg.get_precision() == 'fp32'
            w = weight.detach().float()
            if kind == '1x1':
                fwd = K.pack_conv1x1_weight(w, dt, cim)
                bwd = K.pack_conv1x1_weight(w.reshape(w.shape[0], -1).t().contiguous(), dt, cim)
            else:
                fwd = K.pack_conv_weight(w.contiguous(), dt, cim)
                bwd = K.pack_conv_weight(w.flip(2, 3).transpose(0, 1).contiguous(), dt, cim)
        hit = _pack_cache[key] = (fwd, bwd)
    return hit


def wgrad(g, x, taps, cout, cin):
    """Weight gradient [cout, cin, k, k] of a shared-weight convolution from NHWC g (output gradient) and x (input).  bf16 storage:
    one ood_conv_wgrad call.  fp32 storage (the parity mode): the tensor cores take bf16 operands, so each fp32 operand is split
    into a bf16 head and a bf16 remainder and the three significant products are accumulated (hi*hi + hi*lo + lo*hi: ~2^-16
    relative, against 2^-9 for a plain cast)."""
    g, x = g.contiguous(), x.contiguous()
    if g.dtype == torch.bfloat16:
        return K.conv_wgrad(g, x.to(torch.bfloat16), taps, cout, cin)
    gh, xh = g.to(torch.bfloat16), x.to(torch.bfloat16)
    gl, xl = (g - gh.float()).to(torch.bfloat16), (x - xh.float()).to(torch.bfloat16)
    return K.conv_wgrad(gh, xh, taps, cout, cin) + K.conv_wgrad(gh, xl, taps, cout, cin) + K.conv_wgrad(gl, xh, taps, cout, cin)


def wgrad_ok(g_channels, x_channels, h, w):
    return g_channels % 128 == 0 and x_channels % 64 == 0 and ((w <= 64 and 64 % w == 0 and h % (64 // w) == 0) or (w > 64 and w % 64 == 0))


class _Conv(Function):
    @staticmethod
    def forward(ctx, x, weight, kind):
        fwd, bwd = _packs(weight, kind)
        x = x.contiguous()
        ctx.bwd, ctx.kind, ctx.cin, ctx.wshape = bwd, kind, x.shape[-1], tuple(weight.shape)
        ctx.save_for_backward(x if weight.requires_grad else None)
        y, _ = K.conv3x3(x, fwd, weight.shape[0], transposed=4 if kind == '1x1' else 0, impl=sg._impl())
        return y

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        gw = None
        if ctx.needs_input_grad[1]:
            x, = ctx.saved_tensors
            co, ci = ctx.wshape[:2]
            if not wgrad_ok(g.shape[-1], x.shape[-1], x.shape[1], x.shape[2]):
                raise NotImplementedError(f'ood_gan_inversion_b200: conv weight gradient outside the tcgen05 kernel envelope (Co {g.shape[-1]}, Ci {x.shape[-1]}, '
                                          f'{x.shape[1]}x{x.shape[2]}): needs Co % 128 == 0, Ci % 64 == 0 and 64-pixel patches')
            gw = wgrad(g, x, 1 if ctx.kind == '1x1' else 9, co, ci)
        gx = None
        if ctx.needs_input_grad[0]:
            gx, _ = K.conv3x3(g, ctx.bwd, ctx.cin, transposed=4 if ctx.kind == '1x1' else 0, impl=sg._impl())
        return gx, gw, None


def conv(x, weight, kind='3x3'):
    """NHWC x, Conv2d weight [Co,Ci,3,3] (pad 1, no bias) or [Co,Ci,1,1]; differentiable in x."""
    return _Conv.apply(x, weight, kind)


class _PReLU(Function):
    @staticmethod
    def forward(ctx, x, slope):
        x = x.contiguous()
        ctx.save_for_backward(x, slope)
        return K.prelu(x, slope.detach())

    @staticmethod
    def backward(ctx, g):
        x, slope = ctx.saved_tensors
        g = g.contiguous().to(x.dtype)
        gs = None
        if ctx.needs_input_grad[1]:               # d/dslope[c] = sum g * min(x, 0):  min(x, 0) = x - relu(x)
            relu = K.prelu(x, torch.zeros_like(slope, dtype=torch.float32))
            neg = K.affine2(x, relu, None, torch.full((x.shape[0], x.shape[-1]), -1.0, device=x.device))
            gs = K.dot_reduce(g, neg).sum(0).to(slope.dtype)
        return K.prelu(x, slope.detach(), g=g), gs


def prelu(x, slope):
    return _PReLU.apply(x, slope)


class _Axpy(Function):
    """out = x1 + sign * x2 (NHWC, same shape)."""

    @staticmethod
    def forward(ctx, x1, x2, sign):
        ctx.sign = sign
        b, c = x1.shape[0], x1.shape[-1]
        coef = torch.full((b, c), float(sign), device=x1.device, dtype=torch.float32)
        return K.affine2(x1.contiguous(), x2.contiguous(), None, coef)

    @staticmethod
    def backward(ctx, g):
        g2 = None
        if ctx.needs_input_grad[1]:
            g2 = g if ctx.sign == 1 else K.affine2(g.contiguous(), None, torch.full((g.shape[0], g.shape[-1]), float(ctx.sign), device=g.device))
        return (g if ctx.needs_input_grad[0] else None), g2, None


def add(x1, x2):
    return _Axpy.apply(x1, x2, 1)


def sub(x1, x2):
    return _Axpy.apply(x1, x2, -1)


class _Cat(Function):
    @staticmethod
    def forward(ctx, x1, x2):
        b, h, w, c1 = x1.shape
        c2 = x2.shape[-1]
        ctx.c = (c1, c2)
        out = torch.empty(b, h, w, c1 + c2, device=x1.device, dtype=x1.dtype)
        K.affine2(x1.contiguous(), out=out, channels=c1)
        K.affine2(x2.contiguous(), out=out, channels=c2, off_out=c1)
        return out

    @staticmethod
    def backward(ctx, g):
        c1, c2 = ctx.c
        g = g.contiguous()
        g1 = K.affine2(g, channels=c1) if ctx.needs_input_grad[0] else None
        g2 = K.affine2(g, channels=c2, off1=c1) if ctx.needs_input_grad[1] else None
        return g1, g2


def cat(x1, x2):
    """Channel concatenation of two NHWC tensors."""
    return _Cat.apply(x1, x2)


class _Head27(Function):
    """The AlignNet's 2C -> 3 3x3 convolution as a per-pixel projection onto the 27 (tap, colour) weights + nine shifted partial
    sums (ood_tap_sum): NHWC x -> fp32 NCHW [B,3,H,W]."""

    @staticmethod
    def forward(ctx, x, weight):
        x = x.contiguous()
        w27 = torch.zeros(32, weight.shape[1], device=weight.device)
        w27[:27] = weight.detach().float().permute(2, 3, 0, 1).reshape(27, -1)            # row 3*tap + colour
        dt, cim = sg._act_dtype(), sg.get_precision() == 'fp32'
        proj, _ = K.conv3x3(x, K.pack_conv1x1_weight(w27, dt, cim), 32, transposed=4, impl=sg._impl(), out_f32=True)
        ctx.w27, ctx.c, ctx.dt, ctx.cim = w27, x.shape[-1], dt, cim
        ctx.save_for_backward(x if weight.requires_grad else None)
        return K.tap_sum(proj)

    @staticmethod
    def backward(ctx, g):
        gw = None
        if ctx.needs_input_grad[1]:
            # gW27[r, c] = sum_p G[p, r] * x[p, c]: the pixel contraction with the roles swapped (rows = the 2C channels of x, columns =
            # the gathered taps, padded to 64), then back to Conv2d's [3, 2C, 3, 3]: row 3*tap + colour
            x, = ctx.saved_tensors
            if not wgrad_ok(x.shape[-1], 64, x.shape[1], x.shape[2]):
                raise NotImplementedError('ood_gan_inversion_b200: AlignNet head weight gradient outside the tcgen05 kernel envelope')
            g64 = K.tap_gather(g, 64, ctx.dt)
            gm = wgrad(x, g64, 1, x.shape[-1], 64)[:, :27, 0, 0]                         # [2C, 27]
            gw = gm.t().reshape(3, 3, 3, -1).permute(2, 3, 0, 1).contiguous()          # [ky, kx, colour, c] -> [colour, c, ky, kx]
        gx = None
        if ctx.needs_input_grad[0]:
            gp = K.tap_gather(g, 32, ctx.dt)                                            # [B,H,W,32]
            gx, _ = K.conv3x3(gp, K.pack_conv1x1_weight(ctx.w27.t().contiguous(), ctx.dt, ctx.cim), ctx.c, transposed=4, impl=sg._impl())
        return gx, gw


def head27(x, weight):
    return _Head27.apply(x, weight)


class _Shortcut3(Function):
    """1x1 convolution 2C -> 3 (bottleneck shortcut): NHWC x -> fp32 NCHW [B,3,H,W]; differentiable in x and in the weight."""

    @staticmethod
    def forward(ctx, x, weight):
        x = x.contiguous()
        b = x.shape[0]
        wrgb = weight.detach().float().reshape(1, 3, -1).expand(b, -1, -1).contiguous()
        ctx.save_for_backward(x, wrgb)
        return K.torgb(x, wrgb, torch.zeros(3, device=x.device))

    @staticmethod
    def backward(ctx, g):
        x, wrgb = ctx.saved_tensors
        gx, gw = K.torgb_bwd(g.contiguous(), wrgb, x, None)                             # gw [B,3,C]
        return gx, (gw.sum(0).reshape(3, -1, 1, 1) if ctx.needs_input_grad[1] else None)


def shortcut3(x, weight):
    return _Shortcut3.apply(x, weight)


class _BiasAdd(Function):
    """y = x + bias[c] (NHWC)."""

    @staticmethod
    def forward(ctx, x, bias):
        b, c = x.shape[0], x.shape[-1]
        return K.affine2(x.contiguous(), None, None, None, bias.detach().float().reshape(1, -1).expand(b, -1).contiguous())

    @staticmethod
    def backward(ctx, g):
        gb = None
        if ctx.needs_input_grad[1]:
            g = g.contiguous()
            gb = K.in_stats(g)[..., 0].sum(0) * float(g.shape[1] * g.shape[2])
        return (g if ctx.needs_input_grad[0] else None), gb


def bias_add(x, bias):
    return _BiasAdd.apply(x, bias)
