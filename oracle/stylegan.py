"""Oracle (test infrastructure): StyleGAN2 synthesis network, functional and state-dict driven.

Restates src/ops/StyleGAN/model.py of the reference (line numbers cited per function).  The
state-dict keys are the reference's, so one set of weights drives the
reference, this oracle and the CUDA product.
"""
import math

import torch
import torch.nn.functional as F

from .ops import fir_kernel, fused_leaky_relu, upfirdn2d


def channel_map(channel_multiplier=2):
    """src/ops/StyleGAN/model.py:402-412."""
    cm = channel_multiplier
    return {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * cm, 128: 128 * cm, 256: 64 * cm,
            512: 32 * cm, 1024: 16 * cm}


def equal_linear(x, weight, bias, lr_mul=1.0, activation=False):
    """x @ (W * lr_mul/sqrt(in))^T + b*lr_mul, optionally bias+lrelu*sqrt2.  model.py:129-163."""
    w = weight * (lr_mul / math.sqrt(weight.shape[1]))
    if activation:
        return fused_leaky_relu(F.linear(x, w), bias * lr_mul)
    return F.linear(x, w, None if bias is None else bias * lr_mul)


def pixel_norm(x):
    """model.py:11-16."""
    return x * torch.rsqrt(torch.mean(x * x, dim=1, keepdim=True) + 1e-8)


def mapping_network(sd, prefix, z, n_mlp=8, lr_mlp=0.01):
    """PixelNorm + n_mlp activated EqualLinear layers.  model.py:389-400."""
    x = pixel_norm(z)
    for i in range(1, n_mlp + 1):
        x = equal_linear(x, sd[f'{prefix}style.{i}.weight'], sd[f'{prefix}style.{i}.bias'], lr_mlp, True)
    return x


def modulated_conv2d(x, style, weight, mod_weight, mod_bias, demodulate=True, upsample=False,
                     downsample=False, blur_k=None, blur_pad=None):
    """Per-sample modulated / demodulated convolution.  model.py:233-274.

    weight [1,Co,Ci,k,k]; style [B,style_dim].  `upsample`: stride-2 transposed conv then FIR blur
    (pad (1,1), kernel*4); `downsample`: FIR blur then stride-2 conv.
    """
    b, ci, h, w = x.shape
    _, co, _, k, _ = weight.shape
    s = equal_linear(style, mod_weight, mod_bias)                     # [B,Ci]
    wb = (1.0 / math.sqrt(ci * k * k)) * weight * s.reshape(b, 1, ci, 1, 1)  # [B,Co,Ci,k,k]
    if demodulate:
        d = torch.rsqrt(wb.pow(2).sum([2, 3, 4]) + 1e-8)
        wb = wb * d.reshape(b, co, 1, 1, 1)
    if upsample:
        wt = wb.transpose(1, 2).reshape(b * ci, co, k, k)
        y = F.conv_transpose2d(x.reshape(1, b * ci, h, w), wt, stride=2, padding=0, groups=b)
        y = y.reshape(b, co, y.shape[2], y.shape[3])
        return upfirdn2d(y, blur_k, pad=blur_pad)
    if downsample:
        x = upfirdn2d(x, blur_k, pad=blur_pad)
        h, w = x.shape[2:]
        y = F.conv2d(x.reshape(1, b * ci, h, w), wb.reshape(b * co, ci, k, k), stride=2, groups=b)
        return y.reshape(b, co, y.shape[2], y.shape[3])
    y = F.conv2d(x.reshape(1, b * ci, h, w), wb.reshape(b * co, ci, k, k), padding=k // 2, groups=b)
    return y.reshape(b, co, h, w)


def up_blur_pads(kernel_size=3, taps=4, factor=2):
    """Blur pad after the transposed conv.  model.py:199-205 -> (1,1) for k=3, 4 taps."""
    p = (taps - factor) - (kernel_size - 1)
    return ((p + 1) // 2 + factor - 1, p // 2 + 1)


def skip_up_pads(taps=4, factor=2):
    """Upsample pad for the RGB skip.  model.py:38-43 -> (2,1)."""
    p = taps - factor
    return ((p + 1) // 2 + factor - 1, p // 2)


def styled_conv(sd, prefix, x, style, noise, upsample, hook=None):
    """conv -> (+ w_noise*noise) -> + bias -> lrelu*sqrt2.  model.py:343-350, 283-292.

    `noise` None draws N(0,1) of shape [B,1,H,W] exactly as the reference does.  `hook(image,
    noise, noise_weight)` stands for the reference's callback protocol (model.py:288-290): it must
    return the replacement noise tensor.
    """
    blur_k = sd.get(f'{prefix}conv.blur.kernel')
    y = modulated_conv2d(x, style, sd[f'{prefix}conv.weight'], sd[f'{prefix}conv.modulation.weight'],
                         sd[f'{prefix}conv.modulation.bias'], True, upsample, False, blur_k,
                         up_blur_pads() if upsample else None)
    nw = sd[f'{prefix}noise.weight']
    if noise is None:
        noise = y.new_empty(y.shape[0], 1, y.shape[2], y.shape[3]).normal_()
        if hook is not None:
            noise = hook(y, noise, nw)
    y = y + nw * noise
    return fused_leaky_relu(y, sd[f'{prefix}activate.bias'])


def to_rgb(sd, prefix, x, style, skip=None):
    """1x1 modulated conv (no demod) + bias + FIR-upsampled skip.  model.py:363-372."""
    y = modulated_conv2d(x, style, sd[f'{prefix}conv.weight'], sd[f'{prefix}conv.modulation.weight'],
                         sd[f'{prefix}conv.modulation.bias'], demodulate=False)
    y = y + sd[f'{prefix}bias']
    if skip is not None:
        y = y + upfirdn2d(skip, sd[f'{prefix}upsample.kernel'], up=2, pad=skip_up_pads())
    return y


def generator_forward(sd, latent, size, noise=None, randomize_noise=True, cond_layers=None,
                      hook=None, prefix='', return_features=False):
    """Synthesis forward for W+ latents [B, n_latent, 512].  model.py:483-585 with
    input_is_latent=input_is_tensor=True, cond_type='NOISE'.

    noise: list of per-layer tensors (None entries are drawn); if None and not randomize_noise the
    registered buffers noises.noise_i are used (model.py:503-509).
    hook(cond_index, image, noise, noise_weight, style) -> replacement noise, invoked on the
    up-convolutions whose latent index is in `cond_layers` (model.py:558-569).
    """
    log_size = int(math.log2(size))
    num_layers = (log_size - 2) * 2 + 1
    if noise is None:
        noise = [None] * num_layers if randomize_noise else \
            [sd[f'{prefix}noises.noise_{i}'] for i in range(num_layers)]
    b = latent.shape[0]
    x = sd[f'{prefix}input.input'].repeat(b, 1, 1, 1)
    x = styled_conv(sd, f'{prefix}conv1.', x, latent[:, 0], noise[0], False)
    skip = to_rgb(sd, f'{prefix}to_rgb1.', x, latent[:, 1])
    i = 1
    for blk in range(log_size - 2):
        n1, n2 = noise[1 + 2 * blk], noise[2 + 2 * blk]
        if cond_layers is not None and i in cond_layers and hook is not None:
            ci = cond_layers.index(i)
            style_i = latent[:, i]
            x = styled_conv(sd, f'{prefix}convs.{2 * blk}.', x, style_i, None, True,
                            hook=lambda img, nz, nw, ci=ci, st=style_i: hook(ci, img, nz, nw, st))
        else:
            x = styled_conv(sd, f'{prefix}convs.{2 * blk}.', x, latent[:, i], n1, True)
        x = styled_conv(sd, f'{prefix}convs.{2 * blk + 1}.', x, latent[:, i + 1], n2, False)
        skip = to_rgb(sd, f'{prefix}to_rgbs.{blk}.', x, latent[:, i + 2], skip)
        i += 2
    return (skip, x) if return_features else skip


def synthetic_generator_state(size, style_dim=512, n_mlp=8, channel_multiplier=2, seed=0,
                              rgb_gain=0.1):
    """Random-init state dict with the reference's key names and init distributions
    (model.py:129-141, 219-223, 281, 298-299, 360, 430-433), then the zero-init parameters that
    would hide bugs are made non-zero as SURVEY.md section 8(d) prescribes: noise.weight=0.1,
    activate.bias / to_rgb bias ~ 0.1*N(0,1); to_rgb conv weights are scaled by `rgb_gain` so the
    synthetic image lands near [-1,1].
    """
    g = torch.Generator().manual_seed(seed)

    def rn(*shape):
        return torch.randn(*shape, generator=g)

    ch = channel_map(channel_multiplier)
    sd = {}
    for i in range(1, n_mlp + 1):
        sd[f'style.{i}.weight'] = rn(style_dim, style_dim) / 0.01
        sd[f'style.{i}.bias'] = torch.zeros(style_dim)
    sd['input.input'] = rn(1, ch[4], 4, 4)

    def conv_keys(p, cin, cout, k, up):
        sd[f'{p}conv.weight'] = rn(1, cout, cin, k, k)
        if up:
            sd[f'{p}conv.blur.kernel'] = fir_kernel([1, 3, 3, 1], 4.0)
        sd[f'{p}conv.modulation.weight'] = rn(cin, style_dim)
        sd[f'{p}conv.modulation.bias'] = torch.ones(cin)

    def styled(p, cin, cout, up):
        conv_keys(p, cin, cout, 3, up)
        sd[f'{p}noise.weight'] = torch.full((1,), 0.1)
        sd[f'{p}activate.bias'] = 0.1 * rn(cout)

    def rgb(p, cin, up):
        sd[f'{p}bias'] = 0.1 * rn(1, 3, 1, 1)
        if up:
            sd[f'{p}upsample.kernel'] = fir_kernel([1, 3, 3, 1], 4.0)
        conv_keys(p, cin, 3, 1, False)
        sd[f'{p}conv.weight'] *= rgb_gain

    styled('conv1.', ch[4], ch[4], False)
    rgb('to_rgb1.', ch[4], False)
    log_size = int(math.log2(size))
    cin = ch[4]
    for j, r in enumerate(range(3, log_size + 1)):
        cout = ch[2 ** r]
        styled(f'convs.{2 * j}.', cin, cout, True)
        styled(f'convs.{2 * j + 1}.', cout, cout, False)
        rgb(f'to_rgbs.{j}.', cout, True)
        cin = cout
    for li in range((log_size - 2) * 2 + 1):
        res = (li + 5) // 2
        sd[f'noises.noise_{li}'] = rn(1, 1, 2 ** res, 2 ** res)
    return sd
