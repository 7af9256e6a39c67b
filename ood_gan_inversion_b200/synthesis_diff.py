"""Layer-wise differentiable synthesis: the generator as a graph of per-layer autograd Functions, used when the forward has
something BETWEEN the layers to differentiate through -- the alignment callback of the OOD arch
(src/archs/OOD_faceGAN_e4e_arch.py:224-242 inside src/ops/StyleGAN/model.py:558-571), `return_features`, or a foreign hook.

synthesis_grad.SynthesisFn (one Function for the whole plain synthesis, BASELINE config 4) stays the fast route when there is
nothing in between.  Here every StyledConv is split where the callback cuts it (model.py:277-292):

    img = ModConv(x, w_j)            d * conv(s * x) (+ FIR blur for the up-sampling layers), before noise / bias / activation
    img = callback(img)              optional: aligned features (samm.SPM_Warp.forward_nhwc_diff), differentiable
    y   = NoiseAct(img)              lrelu(img + noise_w * noise + bias) * sqrt2
    rgb = ToRGB(y, w_k, skip)

Backward kernels are the ones of synthesis_grad (act_bwd, the data-gradient convolutions, blur adjoint, dot_reduce, torgb_bwd);
weights are frozen (fix list options/train/E4E_Face.yml:123-125), gradients flow to the W+ latents and through the features.
"""
import torch
from torch.autograd import Function

from . import kernels as K
from . import stylegan as sg
from .synthesis_grad import _bwd_weights, _check


def _style_grad(conv, gs):
    """dL/dstyle-vector [B,D] from dL/ds [B,Ci_p] through the modulation EqualLinear (model.py:129-163)."""
    _, _, mw, _ = conv.packed()
    return (gs @ mw) * conv.modulation.scale


class _ModConv(Function):
    """img = d * conv(s * x) with (s, d) from the style vector; x: UNscaled NHWC activations of the previous layer."""

    @staticmethod
    def forward(ctx, x, style, m):
        conv = m.conv
        x = x.contiguous()
        style = style.detach().float().contiguous()
        s, d = conv.coeffs(style)
        xs = K.nhwc_scale(x, s)
        wp, _, _, _ = conv.packed()
        if conv.upsample:
            t = conv.conv_transposed(xs)
            img, _, _ = K.blur_act(t, conv.blur.taps, d=d, act=False, want_img=True)
        else:
            img, _ = K.conv3x3(xs, wp, conv.cout_p, impl=sg._impl(), d=d)
        ctx.m = m
        ctx.save_for_backward(x, s, d, img)
        return img

    @staticmethod
    def backward(ctx, g):
        x, s, d, img = ctx.saved_tensors
        conv = ctx.m.conv
        g = g.contiguous().to(img.dtype)
        impl = sg._impl()
        g_acc = K.nhwc_scale(g, d)                                  # dL/d(raw accumulator) = g * d
        gd = K.dot_reduce(g, img) / d                               # dL/dd = sum_pix g * acc, acc = img / d
        wd = _bwd_weights(conv)
        if conv.upsample:
            g_t, _, _ = K.blur_act(g_acc, list(reversed(conv.blur.taps)), act=False, want_img=True, pad=(2, 2))
            gxs, gx = K.conv3x3(g_t, wd, conv.cin_p, transposed=2, impl=impl, s_next=s, want_y=True, want_ys=True)
        else:
            gxs, gx = K.conv3x3(g_acc, wd, conv.cin_p, impl=impl, s_next=s, want_y=True, want_ys=True)
        _, wsq, _, _ = conv.packed()
        gs = K.dot_reduce(gxs, x) - s * ((gd * d.pow(3)) @ wsq)
        return gx, _style_grad(conv, gs), None


class _NoiseAct(Function):
    """y = lrelu(img + noise_w * noise + bias) * sqrt2 (model.py:277-292, fused_act.py:92-96)."""

    @staticmethod
    def forward(ctx, img, noise, m):
        nw = m.noise.weight.detach().float()
        bias = sg._pad_dim(m.activate.bias.detach().float(), 0, img.shape[-1])
        y, _ = K.noise_act(img.contiguous(), noise, nw, bias, None, True, False)
        ctx.save_for_backward(y, noise, nw, bias)
        return y

    @staticmethod
    def backward(ctx, g):
        y, noise, nw, bias = ctx.saved_tensors
        ones = torch.ones(y.shape[0], y.shape[-1], device=y.device, dtype=torch.float32)
        g_img, _ = K.act_bwd(g.contiguous().to(y.dtype), y, ones, bias, noise, nw)
        return g_img, None, None


class _ToRGB(Function):
    """rgb = conv1x1(y; w * s) + bias + up2(skip) (model.py:353-372); differentiable in y, the style vector and skip."""

    @staticmethod
    def forward(ctx, y, style, skip, tr):
        wp, _, _, _ = tr.conv.packed()
        s, _ = tr.conv.coeffs(style.detach().float().contiguous())
        wrgb = K.torgb_weight(wp, s, tr.conv.scale)
        y = y.contiguous()
        ctx.tr, ctx.has_skip = tr, skip is not None
        ctx.save_for_backward(y, wrgb)
        return K.torgb(y, wrgb, tr.bias.detach().float().reshape(3).contiguous(), None if skip is None else skip.float().contiguous(),
                       tr.taps_up if skip is not None else None)

    @staticmethod
    def backward(ctx, g):
        y, wrgb = ctx.saved_tensors
        tr = ctx.tr
        g = g.detach().float().contiguous()
        gy, g_wrgb = K.torgb_bwd(g, wrgb, y, None)
        wp, _, _, _ = tr.conv.packed()
        gs = (g_wrgb * wp.unsqueeze(0)).sum(1) * tr.conv.scale
        g_skip = None
        if ctx.has_skip and ctx.needs_input_grad[2]:
            k2 = tr.upsample.kernel.detach().float()
            g_skip = K.upfirdn2d_nchw(g, torch.flip(k2, [0, 1]), 1, 1, 2, 2, 1, 1, 1, 1)       # adjoint of up=2, pad (2,1)
        return gy, _style_grad(tr.conv, gs), g_skip, None


def synthesis(gen, latent, noise, hooks=None, return_features=False):
    """Differentiable W+ synthesis with per-layer hooks.

    latent [B, n_latent, D] (requires grad or not); noise: list of fp32 [B|1,1,R,R], one per layer; hooks: {layer j: fn(img_nhwc)
    -> img_nhwc} applied between the convolution and the noise injection of layer j (j = 1 + 2*blk for the up-sampling layers).
    Returns (image NCHW fp32, last features NHWC or None)."""
    _check(gen)
    hooks = hooks or {}
    b = latent.shape[0]
    lat = latent.float()
    layers = [gen.conv1] + list(gen.convs)
    rgbs = [gen.to_rgb1] + list(gen.to_rgbs)
    n_blocks = gen.log_size - 2

    def layer(j, x, idx):
        m = layers[j]
        img = _ModConv.apply(x, lat[:, idx], m)
        if j in hooks:
            img = hooks[j](img)
        return _NoiseAct.apply(img, noise[j], m)

    x_const = sg._to_nhwc(gen.input.input.detach(), None, gen.conv1.conv.cin_p, batch=b)
    y = layer(0, x_const, 0)
    skip = _ToRGB.apply(y, lat[:, 1], None, rgbs[0])
    i = 1
    for blk in range(n_blocks):
        y = layer(1 + 2 * blk, y, i)
        y = layer(2 + 2 * blk, y, i + 1)
        skip = _ToRGB.apply(y, lat[:, i + 2], skip, rgbs[1 + blk])
        i += 2
    return skip, (y if return_features else None)
