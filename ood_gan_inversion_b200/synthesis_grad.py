"""Differentiable synthesis for optimisation-based W+ inversion (BASELINE config 4): forward + hand-written backward of the
NHWC kernel pipeline as one autograd.Function whose only differentiable input is the W+ latent tensor.

Reference behaviour: autograd through Generator.forward (src/ops/StyleGAN/model.py:483-585) with the generator weights
frozen (fix list options/train/E4E_Face.yml:123-125) and the latents (or `delta_latent`, src/archs/OOD_faceGAN_e4e_arch.py:
126-129) as leaves.  With shared weights the backward is one data-gradient convolution per layer (same implicit-GEMM
kernel, transposed / flipped weights; the stride-2 transposed conv's gradient is the stride-2 gather form) plus
reductions for the style and demodulation gradients -- no weight-gradient GEMM, no saved
per-sample weights, activations saved once in the storage type.
"""
import os

import torch

from . import kernels as K
from . import stylegan as sg

_FUSED_BWD = os.environ.get('OOD_FUSED_BWD', '1') != '0'      # A/B switch: one fused pass per layer (ood_act_bwd_fused) vs torgb_bwd + act_bwd + dot_reduce


def _bwd_weights(conv):
    """Packed data-gradient weights of a ModulatedConv2d, cached on the module."""
    key = (sg.get_precision(), conv.weight.device, conv.weight._version, conv.weight.data_ptr())
    hit = conv._cache.get('bwd')
    if hit is None or hit[0] != key:
        with torch.no_grad():
            w = conv.weight.detach()[0].float()                       # [Co, Ci, 3, 3]
            wd = w.transpose(0, 1) if conv.upsample else w.flip(2, 3).transpose(0, 1)
            pk = K.pack_conv_weight(wd.contiguous(), sg._act_dtype(), ci_major=(sg.get_precision() == 'fp32'))
        conv._cache['bwd'] = (key, pk)
        hit = conv._cache['bwd']
    return hit[1]


def _check(gen):
    g = sg._granule()
    for m in [gen.conv1] + list(gen.convs):
        if m.conv.in_channel % g or m.conv.out_channel % g:
            raise NotImplementedError(f'ood_gan_inversion_b200: the differentiable path needs channel counts that are multiples of {g}')


class SynthesisFn(torch.autograd.Function):
    """image = synthesis(latent[B, n_latent, D]); noise tensors and the generator are non-differentiable inputs."""

    @staticmethod
    def forward(ctx, latent, gen, noise):
        _check(gen)
        lat = latent.detach().float().contiguous()
        b = lat.shape[0]
        n_blocks = gen.log_size - 2
        layers = [gen.conv1] + list(gen.convs)
        sd = [m.conv.coeffs(lat[:, 0 if j == 0 else j]) for j, m in enumerate(layers)]      # latent index of layer j is j (0, 1..2n)
        rgbs = [gen.to_rgb1] + list(gen.to_rgbs)
        saved = dict(sd=sd, noise=noise, y=[], wrgb=[], s_rgb=[])

        def rgb_weight(tr, idx):
            wp, _, _, _ = tr.conv.packed()
            s, _ = tr.conv.coeffs(lat[:, idx])
            wrgb = K.torgb_weight(wp, s, tr.conv.scale)
            saved['wrgb'].append(wrgb)
            saved['s_rgb'].append(s)
            return wrgb

        x_const = sg._to_nhwc(gen.input.input.detach(), None, gen.conv1.conv.cin_p, batch=b)    # unscaled, for the style gradient
        xs = K.nhwc_scale(x_const, sd[0][0])
        y, ys = gen.conv1.run_nhwc(xs, None, noise[0], s_next=sd[1][0] if n_blocks else None, want_y=True,
                                   want_ys=n_blocks > 0, d=sd[0][1])
        saved['y'].append(y)
        skip = K.torgb(y, rgb_weight(rgbs[0], 1), rgbs[0].bias.detach().float().reshape(3).contiguous())
        for blk in range(n_blocks):
            c1, c2, tr = layers[1 + 2 * blk], layers[2 + 2 * blk], rgbs[1 + blk]
            y1, y1s = c1.run_nhwc(ys, None, noise[1 + 2 * blk], s_next=sd[2 + 2 * blk][0], want_y=True, want_ys=True, d=sd[1 + 2 * blk][1])
            last = blk == n_blocks - 1
            y, ys = c2.run_nhwc(y1s, None, noise[2 + 2 * blk], s_next=None if last else sd[3 + 2 * blk][0], want_y=True,
                                want_ys=not last, d=sd[2 + 2 * blk][1])
            saved['y'] += [y1, y]
            skip = K.torgb(y, rgb_weight(tr, 3 + 2 * blk), tr.bias.detach().float().reshape(3).contiguous(), skip, tr.taps_up)
        saved['x_const'] = x_const
        ctx.gen, ctx.saved, ctx.lat_shape = gen, saved, latent.shape
        ctx.lat_dtype = latent.dtype
        return skip

    @staticmethod
    def backward(ctx, g_image):
        if ctx.saved is None:
            raise RuntimeError('ood_gan_inversion_b200: SynthesisFn.backward ran twice: the saved activations are freed after the first '
                               'backward (retain_graph / double backward are not supported on this path; the layer-wise graph of '
                               'synthesis_diff supports retain_graph)')
        gen, S = ctx.gen, ctx.saved
        layers = [gen.conv1] + list(gen.convs)
        rgbs = [gen.to_rgb1] + list(gen.to_rgbs)
        n_blocks = gen.log_size - 2
        b = g_image.shape[0]
        dev = g_image.device
        g_lat = torch.zeros(ctx.lat_shape, device=dev, dtype=torch.float32)
        impl = sg._impl()

        def add_style_grad(idx, gs, conv):
            _, _, mw, _ = conv.packed()                                   # [Ci_p, D] fp32 (EqualLinear scale applied below)
            g_lat[:, idx] += (gs @ mw) * conv.modulation.scale

        def conv_bwd(j, gy, x_in):
            """Layer j (StyledConv): gy = dL/dy_j (NHWC).  Returns dL/dx_in (NHWC, unscaled input) and accumulates dL/dlatent."""
            m = layers[j]
            conv = m.conv
            s, d = S['sd'][j]
            nw = m.noise.weight.detach().float()
            bias = m.activate.bias.detach().float().contiguous()
            g_pre, gd = K.act_bwd(gy.contiguous(), S['y'][j], d, bias, S['noise'][j], nw)
            wd = _bwd_weights(conv)
            if conv.upsample:
                g_t, _, _ = K.blur_act(g_pre, list(reversed(conv.blur.taps)), act=False, want_img=True, pad=(2, 2))
                gxs, gx = K.conv3x3(g_t, wd, conv.cin_p, transposed=2, impl=impl, s_next=s, want_y=True, want_ys=True)
            else:
                gxs, gx = K.conv3x3(g_pre, wd, conv.cin_p, impl=impl, s_next=s, want_y=True, want_ys=True)
            _, wsq, _, _ = conv.packed()
            gs = K.dot_reduce(gxs, x_in) - s * ((gd * d.pow(3)) @ wsq)
            add_style_grad(j, gs, conv)
            return gx

        def rgb_bwd(r, g_rgb, y, g_in):
            tr = rgbs[r]
            gy, g_wrgb = K.torgb_bwd(g_rgb, S['wrgb'][r], y, g_in)
            wp, _, _, _ = tr.conv.packed()                                # [3, Ci_p]
            gs = (g_wrgb * wp.unsqueeze(0)).sum(1) * tr.conv.scale
            add_style_grad(1 if r == 0 else 1 + 2 * r, gs, tr.conv)
            return gy

        if _FUSED_BWD:
            # One streaming pass per layer (ood_act_bwd_fused) instead of torgb_bwd (2 kernels) + act_bwd + dot_reduce: the data-gradient
            # convolution of layer j writes ONE tensor, the unscaled dL/d(s*x); the pass of the layer below applies s on load, adds the ToRGB
            # gradient of its level, gates, and returns the three reductions -- among them sum_pix gxs_j * x_j, layer j's style gradient,
            # because x_j IS the saved output of the layer below.  10 -> 4 tensor passes per ToRGB layer, 7 -> 4 for the others.
            def layer_bwd(j, g_in, g_scale, rgb, pending):
                m = layers[j]
                conv = m.conv
                s, d = S['sd'][j]
                nw = m.noise.weight.detach().float()
                bias = m.activate.bias.detach().float().contiguous()
                g_pre, gd, dot, g_w = K.act_bwd_fused(g_in, g_scale, rgb, S['y'][j], d, bias, S['noise'][j], nw)
                if pending is not None:                                   # the layer above: its input was this layer's output
                    finish_style(pending, dot)
                wd = _bwd_weights(conv)
                if conv.upsample:
                    g_t, _, _ = K.blur_act(g_pre, list(reversed(conv.blur.taps)), act=False, want_img=True, pad=(2, 2))
                    gxs, _ = K.conv3x3(g_t, wd, conv.cin_p, transposed=2, impl=impl, want_y=True, want_ys=False)
                else:
                    gxs, _ = K.conv3x3(g_pre, wd, conv.cin_p, impl=impl, want_y=True, want_ys=False)
                return gxs, g_w, (j, s, d, gd)

            def finish_style(pending, dot):
                j, s, d, gd = pending
                conv = layers[j].conv
                _, wsq, _, _ = conv.packed()
                add_style_grad(j, dot - s * ((gd * d.pow(3)) @ wsq), conv)

            def rgb_style(r, g_w):
                tr = rgbs[r]
                wp, _, _, _ = tr.conv.packed()                                # [3, Ci_p]
                add_style_grad(1 if r == 0 else 1 + 2 * r, (g_w * wp.unsqueeze(0)).sum(1) * tr.conv.scale, tr.conv)

            g_skip = g_image.detach().float().contiguous()
            gxs, pending = None, None
            for blk in reversed(range(n_blocks)):
                j1, j2 = 1 + 2 * blk, 2 + 2 * blk
                tr = rgbs[1 + blk]
                gxs, g_w, pending = layer_bwd(j2, gxs, None if pending is None else pending[1], (g_skip, S['wrgb'][1 + blk]), pending)
                rgb_style(1 + blk, g_w)
                k2 = tr.upsample.kernel.detach().float()
                g_skip = K.upfirdn2d_nchw(g_skip, torch.flip(k2, [0, 1]), 1, 1, 2, 2, 1, 1, 1, 1)     # adjoint of up=2, pad (2,1)
                gxs, _, pending = layer_bwd(j1, gxs, pending[1], None, pending)
            gxs, g_w, pending = layer_bwd(0, gxs, None if pending is None else pending[1], (g_skip, S['wrgb'][0]), pending)
            rgb_style(0, g_w)
            finish_style(pending, K.dot_reduce(gxs, S['x_const']))           # the constant input has no layer below it
            ctx.saved = None
            return g_lat.to(ctx.lat_dtype), None, None

        g_skip = g_image.detach().float().contiguous()
        gy_next = None
        for blk in reversed(range(n_blocks)):
            j1, j2 = 1 + 2 * blk, 2 + 2 * blk
            tr = rgbs[1 + blk]
            gy2 = rgb_bwd(1 + blk, g_skip, S['y'][j2], gy_next)
            k2 = tr.upsample.kernel.detach().float()
            g_skip = K.upfirdn2d_nchw(g_skip, torch.flip(k2, [0, 1]), 1, 1, 2, 2, 1, 1, 1, 1)     # adjoint of up=2, pad (2,1)
            gy1 = conv_bwd(j2, gy2, S['y'][j1])
            gy_next = conv_bwd(j1, gy1, S['y'][j1 - 1])
        gy0 = rgb_bwd(0, g_skip, S['y'][0], gy_next)
        conv_bwd(0, gy0, S['x_const'])
        ctx.saved = None
        return g_lat.to(ctx.lat_dtype), None, None


def synthesis(gen, latent, noise):
    """Differentiable w.r.t. `latent`; `noise`: list of fp32 [B|1,1,R,R] tensors, one per layer."""
    return SynthesisFn.apply(latent, gen, noise)
