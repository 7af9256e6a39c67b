"""Drop-in for the generator side of the reference's `src/ops/StyleGAN/model.py`.

Same class names, constructor arguments, `forward` keywords (incl. the callback / noise / index / style protocol)
and state-dict keys, so rosinality `g_ema` checkpoints load unchanged.  The arithmetic runs
on the sm_100a kernels behind the C ABI:

  * activations live in NHWC (bf16 on the tcgen05 path, fp32 on the SIMT parity path);
  * (W*s*d) (*) x == d . (W (*) (s . x)): weights stay shared, the style scale of the NEXT convolution is
    written by the producing kernel's epilogue, demodulation / noise / bias / leaky-ReLU are fused after the GEMM
    (stride-1 conv) or after the FIR blur (transposed conv);
  * the RGB skip is accumulated in fp32 NCHW by the fused ToRGB + up-FIR kernel.

`Generator.forward` runs that pipeline end to end; the individual modules (`ModulatedConv2d`, `StyledConv`,
`ToRGB`, `Blur`, ...) keep the reference's NCHW-in / NCHW-out contract for standalone use.
Forward only for the module path in this round (the autograd Functions cover upfirdn2d / fused_leaky_relu).
"""
import math
import os
import random

import torch
from torch import nn
from torch.nn import functional as F

from . import kernels as K
from .op import FusedLeakyReLU, fused_leaky_relu, upfirdn2d

_PRECISION = os.environ.get('OOD_B200_PRECISION', 'bf16')
# up-sampling layers with at most this many output channels use the fused-phase transposed convolution (conv3x3 form 5): the
# zero-padded weight blocks cost 1.8x the MACs, which only the HBM / per-tile-overhead bound small-channel layers can afford
_FUSED_T_MAX_CO = int(os.environ.get('OOD_FUSED_T_MAX_CO', 64))
_FUSE_RGB_ROWS = int(os.environ.get('OOD_FUSE_RGB_ROWS', '0'))  # ToRGB in the row-sliding kernel's epilogue (512 / 1024 px layers): measured, OFF (see _synthesis)
_FUSE_RGB_TC = os.environ.get('OOD_FUSE_RGB_TC', '1') != '0'    # ToRGB in the generic tiles' epilogue (128 / 256 channels); A/B switch
_CONVT_ROWS = os.environ.get('OOD_CONVT_ROWS', '1') != '0'      # the row-streaming transposed kernel for the 64 -> 32 layer (csrc/convt_rows.cu)


def set_precision(p):
    """'bf16': tcgen05 tensor cores, bf16 activations.  'fp32': fp32 SIMT parity mode."""
    global _PRECISION
    if p not in ('bf16', 'fp32'):
        raise ValueError(p)
    _PRECISION = p


def get_precision():
    return _PRECISION


def _act_dtype():
    return torch.bfloat16 if _PRECISION == 'bf16' else torch.float32


def _impl():
    return 0 if _PRECISION == 'bf16' else 1


def _granule():
    """Channel granule of the active conv kernel (tcgen05: 32, SIMT: 16).  Channel counts that are not multiples are
    zero-padded on the weight side (compatibility path; every StyleGAN2 channel count is already a multiple of 32)."""
    return 32 if _PRECISION == 'bf16' else 16


def _round_up(v, g):
    return (v + g - 1) // g * g


def _pad_dim(t, dim, size):
    if t.shape[dim] == size:
        return t
    pad = [0, 0] * (t.dim() - dim - 1) + [0, size - t.shape[dim]]
    return F.pad(t, pad)


def _to_nhwc(x, scale, cin_p, batch=None):
    """NCHW fp32 -> NHWC storage type with `cin_p` (zero-padded) channels, times scale[b, c]."""
    if x.shape[1] != cin_p:
        x = _pad_dim(x.float(), 1, cin_p)
    return K.nchw_to_nhwc(x, scale, _act_dtype(), batch=batch)


def _to_nchw(y, channels):
    if y.shape[-1] != channels:
        y = y[..., :channels].contiguous()
    return K.nhwc_to_nchw(y)


class PixelNorm(nn.Module):
    """model.py:11-16"""

    def forward(self, input):
        return input * torch.rsqrt(torch.mean(input ** 2, dim=1, keepdim=True) + 1e-8)


def make_kernel(k):
    """model.py:19-27"""
    k = torch.tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = k[None, :] * k[:, None]
    k /= k.sum()
    return k


class Upsample(nn.Module):
    """model.py:30-47"""

    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        self.register_buffer('kernel', make_kernel(kernel) * (factor ** 2))
        p = self.kernel.shape[0] - factor
        self.pad = ((p + 1) // 2 + factor - 1, p // 2)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=self.factor, down=1, pad=self.pad)


class Downsample(nn.Module):
    """model.py:50-68"""

    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        self.register_buffer('kernel', make_kernel(kernel))
        p = self.kernel.shape[0] - factor
        self.pad = ((p + 1) // 2, p // 2)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=1, down=self.factor, pad=self.pad)


class Blur(nn.Module):
    """model.py:71-88"""

    def __init__(self, kernel, pad, upsample_factor=1):
        super().__init__()
        k2 = make_kernel(kernel)
        if upsample_factor > 1:
            k2 = k2 * (upsample_factor ** 2)
        self.register_buffer('kernel', k2)
        self.pad = pad
        k1 = [float(v) for v in kernel]
        self.taps = [v / sum(k1) * upsample_factor for v in k1] if len(k1) == 4 else None   # separable 1-D form

    def forward(self, input):
        return upfirdn2d(input, self.kernel, pad=self.pad)


class EqualLinear(nn.Module):
    """model.py:129-163 (tiny GEMM: stays a library call; the activation is the fused kernel)."""

    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, lr_mul=1, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        self.bias = nn.Parameter(torch.zeros(out_dim).fill_(bias_init)) if bias else None
        self.activation = activation
        self.scale = (1 / math.sqrt(in_dim)) * lr_mul
        self.lr_mul = lr_mul

    def forward(self, input):
        if self.activation:
            out = F.linear(input, self.weight * self.scale)
            return fused_leaky_relu(out, self.bias * self.lr_mul)
        return F.linear(input, self.weight * self.scale, bias=self.bias * self.lr_mul)

    def __repr__(self):
        return f'{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]})'


class ModulatedConv2d(nn.Module):
    """model.py:178-274.  3x3 (plain / upsample / downsample) and the 1x1 -> 3 ToRGB form."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True, upsample=False,
                 downsample=False, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        self.eps = 1e-8
        self.kernel_size = kernel_size
        self.in_channel = in_channel
        self.out_channel = out_channel
        self.upsample = upsample
        self.downsample = downsample
        if upsample:
            factor = 2
            p = (len(blur_kernel) - factor) - (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2 + factor - 1, p // 2 + 1), upsample_factor=factor)
        if downsample:
            factor = 2
            p = (len(blur_kernel) - factor) + (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2, p // 2))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.padding = kernel_size // 2
        self.weight = nn.Parameter(torch.randn(1, out_channel, in_channel, kernel_size, kernel_size))
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1)
        self.demodulate = demodulate
        self._cache = {}

    def __repr__(self):
        return (f'{self.__class__.__name__}({self.in_channel}, {self.out_channel}, {self.kernel_size}, '
                f'upsample={self.upsample}, downsample={self.downsample})')

    # ---- kernel-side state -------------------------------------------------------------------------------
    @property
    def cin_p(self):
        return _round_up(self.in_channel, _granule())

    @property
    def cout_p(self):
        return 3 if self.kernel_size == 1 else _round_up(self.out_channel, _granule())

    def packed(self):
        """(packed conv weight for the active precision, Wsq[Co,Ci] or None, fp32 mod weight, fp32 mod bias); channel
        dimensions zero-padded to the kernel granule."""
        key = (_PRECISION, self.weight.device, self.weight._version, self.weight.data_ptr(), self.modulation.weight._version,
               self.modulation.weight.data_ptr(), self.modulation.bias._version, self.modulation.bias.data_ptr())
        hit = self._cache.get('k')
        if hit is None or hit[0] != key:
            with torch.no_grad():
                cin_p, cout_p = self.cin_p, self.cout_p
                w = self.weight.detach()[0].float()
                w = _pad_dim(_pad_dim(w, 1, cin_p), 0, cout_p).contiguous()
                if self.kernel_size == 1:
                    wp = w.reshape(cout_p, cin_p).contiguous()
                else:
                    wp = K.pack_conv_weight(w, _act_dtype(), ci_major=(_PRECISION == 'fp32'))
                wsq = K.weight_sumsq(w) if self.demodulate else None
                mw = _pad_dim(self.modulation.weight.detach().float(), 0, cin_p).contiguous()
                mb = _pad_dim(self.modulation.bias.detach().float() * self.modulation.lr_mul, 0, cin_p).contiguous()
                # small-channel up-sampling layers (512 / 1024 px): fused-phase transposed form (conv3x3 transposed=5)
                # (the 64 -> 32 layer has its own row-streaming kernel behind form 1: csrc/convt_rows.cu)
                self._cache['fused_t'] = K.pack_convt_fused(wp) if (self.upsample and self.kernel_size == 3 and _PRECISION == 'bf16'
                                                                      and cout_p <= _FUSED_T_MAX_CO and cin_p % 32 == 0
                                                                      and not (_CONVT_ROWS and cin_p == 64 and cout_p == 32)) else None
            self._cache['k'] = (key, (wp, wsq, mw, mb))
            hit = self._cache['k']
        return hit[1]

    def conv_transposed(self, xs):
        """Raw accumulators [B,2h+1,2w+1,Co] of the stride-2 transposed convolution (model.py:246-256) of pre-modulated xs."""
        wp, _, _, _ = self.packed()
        wf = self._cache.get('fused_t')
        if wf is not None and _impl() == 0:
            return K.conv3x3(xs, wf, self.cout_p, transposed=5, impl=0)[0]
        return K.conv3x3(xs, wp, self.cout_p, transposed=True, impl=_impl())[0]

    def coeffs(self, style):
        """style [B,style_dim] fp32 (may be a strided row view of the W+ tensor) -> (s [B,Ci], d [B,Co])."""
        wp, wsq, mw, mb = self.packed()
        if style.dtype != torch.float32 or style.stride(-1) != 1:
            style = style.float().contiguous()
        return K.modulation(style, mw, mb, wsq, self.scale, self.cout_p, want_d=self.kernel_size != 1)

    def forward(self, input, style):
        if self.kernel_size == 1:
            if self.out_channel != 3 or self.demodulate:
                raise NotImplementedError('ood_gan_inversion_b200: the 1x1 modulated conv is implemented for the ToRGB '
                                          'form (3 output channels, demodulate=False) only')
            wp, _, _, _ = self.packed()
            s, _ = self.coeffs(style)
            y = _to_nhwc(input, None, self.cin_p)
            zero = torch.zeros(3, device=input.device)
            return K.torgb(y, K.torgb_weight(wp, s, self.scale), zero)
        if self.kernel_size != 3:
            raise NotImplementedError('ood_gan_inversion_b200: ModulatedConv2d supports kernel_size 3 (and the 1x1 ToRGB form)')
        wp, _, _, _ = self.packed()
        s, d = self.coeffs(style)
        if self.downsample:   # blur, then the stride-2 conv = odd samples of the pad-1 stride-1 conv
            input = self.blur(input)
            xs = _to_nhwc(input, s, self.cin_p)
            y, _ = K.conv3x3(xs, wp, self.cout_p, impl=_impl(), d=d)
            oh, ow = (input.shape[2] - 3) // 2 + 1, (input.shape[3] - 3) // 2 + 1
            return _to_nchw(y, self.out_channel)[:, :, 1::2, 1::2][:, :, :oh, :ow].contiguous()
        xs = _to_nhwc(input, s, self.cin_p)
        if self.upsample:
            t = self.conv_transposed(xs)
            img, _, _ = K.blur_act(t, self.blur.taps, d=d, act=False, want_img=True)
            return _to_nchw(img, self.out_channel)
        y, _ = K.conv3x3(xs, wp, self.cout_p, impl=_impl(), d=d)
        return _to_nchw(y, self.out_channel)


class NoiseInjection(nn.Module):
    """model.py:277-292"""

    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))

    def forward(self, image, noise=None, **kwargs):
        if noise is None:
            batch, _, height, width = image.shape
            noise = image.new_empty(batch, 1, height, width).normal_()
            if kwargs.get('callback', None):
                kwargs.update({'noise_weight': self.weight, 'noise': noise})
                noise = kwargs.get('callback')(image, **kwargs)
        return image + self.weight * noise


class ConstantInput(nn.Module):
    """model.py:295-305"""

    def __init__(self, channel, size=4):
        super().__init__()
        self.input = nn.Parameter(torch.randn(1, channel, size, size))

    def forward(self, input):
        return self.input.repeat(input.shape[0], 1, 1, 1)


class StyledConv(nn.Module):
    """model.py:308-350: conv -> noise -> bias + leaky-ReLU*sqrt2."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=[1, 3, 3, 1],
                 demodulate=True, **kwargs):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample,
                                    blur_kernel=blur_kernel, demodulate=demodulate)
        self.noise = NoiseInjection() if kwargs.get('noiseInjection', True) else (lambda x, **kw: x)
        self.activate = FusedLeakyReLU(out_channel) if kwargs.get('activation', True) else (lambda x: x)

    # ---- NHWC pipeline stage (used by Generator and by forward) --------------------------------------------
    def run_nhwc(self, xs, style, noise, s_next=None, want_y=True, want_ys=False, hook=None, d=None, rgb=None):
        """xs: NHWC activations already carrying this conv's style scale.  noise: fp32 [B|1,1,H,W] (required).
        hook(image_nhwc) -> replacement image (NHWC), applied between the convolution and the noise injection.
        d: demodulation coefficients if the caller already has them (else computed from `style`)."""
        conv = self.conv
        wp, _, _, _ = conv.packed()
        if d is None:
            _, d = conv.coeffs(style)
        has_noise = isinstance(self.noise, NoiseInjection)
        has_act = isinstance(self.activate, FusedLeakyReLU)
        nw = self.noise.weight.detach().float() if has_noise else None
        bias = _pad_dim(self.activate.bias.detach().float(), 0, conv.cout_p) if has_act else None
        if not has_noise:
            noise = None
        if not has_act:
            raise NotImplementedError('ood_gan_inversion_b200: StyledConv(activation=False) is not on the hot path')
        if conv.upsample:
            t = conv.conv_transposed(xs)
            if hook is None:
                _, y, ys = K.blur_act(t, conv.blur.taps, d=d, noise=noise, noise_w=nw, bias=bias, s_next=s_next, act=True,
                                      want_y=want_y, want_ys=want_ys)
                return y, ys
            img, _, _ = K.blur_act(t, conv.blur.taps, d=d, act=False, want_img=True)
            img = hook(img)
            return K.noise_act(img, noise, nw, bias, s_next, want_y, want_ys)
        if hook is None:
            return K.conv3x3(xs, wp, conv.cout_p, impl=_impl(), d=d, noise=noise, noise_w=nw, bias=bias, s_next=s_next,
                             act=True, want_y=want_y, want_ys=want_ys, rgb=rgb)
        img, _ = K.conv3x3(xs, wp, conv.cout_p, impl=_impl(), d=d)
        img = hook(img)
        return K.noise_act(img, noise, nw, bias, s_next, want_y, want_ys)

    def forward(self, input, style, noise=None, **kwargs):
        conv = self.conv
        s, _ = conv.coeffs(style)
        xs = _to_nhwc(input, s, conv.cin_p)
        b, _, h, w = input.shape
        oh, ow = (2 * h, 2 * w) if conv.upsample else (h, w)
        kwargs.update({'style': style})
        hook = None
        if noise is None and isinstance(self.noise, NoiseInjection):
            noise = input.new_empty(b, 1, oh, ow, dtype=torch.float32).normal_()
            if kwargs.get('callback', None):
                hook, noise = _callback_hook(self.noise, noise, kwargs, conv.out_channel)
        elif noise is not None:
            noise = noise.float().contiguous()
            if noise.shape[1] != 1:
                raise NotImplementedError('ood_gan_inversion_b200: explicit noise must be [B|1,1,H,W]')
        y, _ = self.run_nhwc(xs, style, noise, hook=hook)
        return _to_nchw(y, conv.out_channel)


def _callback_hook(noise_mod, noise, kwargs, channels):
    """Bridges the reference's callback protocol (model.py:288-292: noise := callback(image, noise_weight=, noise=,
    style=, index=, ...); image + weight*noise) onto the NHWC pipeline.

    A callback that exposes `aligned_nhwc(image_nhwc, **kwargs)` (this package's own arch) stays in NHWC and the noise
    injection stays fused; any other callable gets the reference contract verbatim (NCHW fp32 in, replacement
    "noise" out), evaluated with torch ops."""
    cb = kwargs['callback']
    kw = dict(kwargs)
    kw.update({'noise_weight': noise_mod.weight, 'noise': noise})
    if hasattr(cb, 'aligned_nhwc'):
        return (lambda img: cb.aligned_nhwc(img, **kw)), noise

    def generic(img):
        cp = img.shape[-1]
        image = _to_nchw(img, channels)
        repl = cb(image, **kw)
        return _to_nhwc(image + noise_mod.weight.detach() * repl, None, cp)
    return generic, None      # the replacement already contains the noise term


class ToRGB(nn.Module):
    """model.py:353-372"""

    def __init__(self, in_channel, style_dim, upsample=True, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        if upsample:
            self.upsample = Upsample(blur_kernel)
        self.conv = ModulatedConv2d(in_channel, 3, 1, style_dim, demodulate=False)
        self.bias = nn.Parameter(torch.zeros(1, 3, 1, 1))
        k1 = [float(v) for v in blur_kernel]
        self.taps_up = [v / sum(k1) * 2 for v in k1]

    def fused_args(self, style, skip):
        """(wrgb, bias, skip, taps) for the conv epilogue that computes this ToRGB in place (ood_conv3x3 rgb_* fields)."""
        wp, _, _, _ = self.conv.packed()
        s, _ = self.conv.coeffs(style)
        return (K.torgb_weight(wp, s, self.conv.scale), self.bias.detach().float().reshape(3).contiguous(),
                None if skip is None else skip.float().contiguous(), self.taps_up)

    def run_nhwc(self, y, style, skip=None):
        """y: UNscaled NHWC activations; skip: NCHW fp32 [B,3,H/2,W/2] or None -> NCHW fp32 [B,3,H,W]."""
        wp, _, _, _ = self.conv.packed()
        s, _ = self.conv.coeffs(style)
        if skip is not None and len(self.taps_up) != 4:
            raise NotImplementedError('ood_gan_inversion_b200: fused ToRGB skip needs a 4-tap blur kernel')
        return K.torgb(y, K.torgb_weight(wp, s, self.conv.scale), self.bias.detach().float().reshape(3).contiguous(),
                       None if skip is None else skip.float().contiguous(), self.taps_up)

    def forward(self, input, style, skip=None):
        return self.run_nhwc(_to_nhwc(input, None, self.conv.cin_p), style, skip)


_WARNED_FROZEN = False


class Generator(nn.Module):
    """model.py:375-585: same constructor / forward contract; the synthesis runs on the NHWC kernel pipeline."""

    def __init__(self, size, style_dim, n_mlp, channel_multiplier=2, blur_kernel=[1, 3, 3, 1], lr_mlp=0.01):
        super().__init__()
        self.size = size
        self.style_dim = style_dim
        layers = [PixelNorm()]
        for _ in range(n_mlp):
            layers.append(EqualLinear(style_dim, style_dim, lr_mul=lr_mlp, activation='fused_lrelu'))
        self.style = nn.Sequential(*layers)
        cm = channel_multiplier
        self.channels = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * cm, 128: 128 * cm, 256: 64 * cm, 512: 32 * cm,
                         1024: 16 * cm}
        self.input = ConstantInput(self.channels[4])
        self.conv1 = StyledConv(self.channels[4], self.channels[4], 3, style_dim, blur_kernel=blur_kernel)
        self.to_rgb1 = ToRGB(self.channels[4], style_dim, upsample=False)
        self.log_size = int(math.log(size, 2))
        self.num_layers = (self.log_size - 2) * 2 + 1
        self.convs = nn.ModuleList()
        self.upsamples = nn.ModuleList()
        self.to_rgbs = nn.ModuleList()
        self.noises = nn.Module()
        in_channel = self.channels[4]
        for layer_idx in range(self.num_layers):
            res = (layer_idx + 5) // 2
            self.noises.register_buffer(f'noise_{layer_idx}', torch.randn(1, 1, 2 ** res, 2 ** res))
        for i in range(3, self.log_size + 1):
            out_channel = self.channels[2 ** i]
            self.convs.append(StyledConv(in_channel, out_channel, 3, style_dim, upsample=True, blur_kernel=blur_kernel))
            self.convs.append(StyledConv(out_channel, out_channel, 3, style_dim, blur_kernel=blur_kernel))
            self.to_rgbs.append(ToRGB(out_channel, style_dim))
            in_channel = out_channel
        self.n_latent = self.log_size * 2 - 2

    def make_noise(self):
        device = self.input.input.device
        noises = [torch.randn(1, 1, 2 ** 2, 2 ** 2, device=device)]
        for i in range(3, self.log_size + 1):
            for _ in range(2):
                noises.append(torch.randn(1, 1, 2 ** i, 2 ** i, device=device))
        return noises

    def mean_latent(self, n_latent):
        latent_in = torch.randn(n_latent, self.style_dim, device=self.input.input.device)
        return self.style(latent_in).mean(0, keepdim=True)

    def get_latent(self, input):
        return self.style(input)

    def forward(self, styles, return_latents=False, return_features=False, inject_index=None, truncation=1,
                truncation_latent=None, input_is_latent=False, input_is_tensor=False, noise=None, randomize_noise=True,
                conditions=None, cond_layers=None, cond_type='SFT', **kwargs):
        if not input_is_latent and not input_is_tensor:
            styles = [self.style(s) for s in styles]
        if noise is None:
            if randomize_noise:
                noise = [None] * self.num_layers
            else:
                noise = [getattr(self.noises, f'noise_{i}') for i in range(self.num_layers)]
        if truncation < 1:
            styles = [truncation_latent + truncation * (s - truncation_latent) for s in styles]
        if not input_is_tensor:
            if len(styles) < 2:
                inject_index = self.n_latent
                latent = styles[0].unsqueeze(1).repeat(1, inject_index, 1) if styles[0].ndim < 3 else styles[0]
            else:
                if inject_index is None:
                    inject_index = random.randint(1, self.n_latent - 1)
                latent = styles[0].unsqueeze(1).repeat(1, inject_index, 1)
                latent2 = styles[1].unsqueeze(1).repeat(1, self.n_latent - inject_index, 1)
                latent = torch.cat([latent, latent2], 1)
        else:
            latent = styles
        if cond_layers is not None and conditions is not None and cond_type != 'NOISE':
            raise NotImplementedError("ood_gan_inversion_b200: only cond_type='NOISE' is implemented (the SFT/ADD/FUSE "
                                      'branches are never exercised by the shipped configs, SURVEY appendix B.5)')
        image, feat = self._synthesis(latent, noise, conditions, cond_layers, return_features, kwargs)
        if return_latents:
            return image, latent
        if return_features:
            return image, feat
        return image, None

    # ---- the hot path ---------------------------------------------------------------------------------------
    def _synthesis(self, latent, noise, conditions, cond_layers, return_features, kwargs):
        if not latent.is_cuda:
            raise RuntimeError('ood_gan_inversion_b200 is CUDA-only: move the generator and latents to a CUDA device')
        has_cond = cond_layers is not None and conditions is not None
        grad_route = torch.is_grad_enabled() and (latent.requires_grad or bool(kwargs.get('differentiable', False)))
        if torch.is_grad_enabled() and not grad_route and any(p.requires_grad for p in self.parameters()):
            global _WARNED_FROZEN
            if not _WARNED_FROZEN:
                _WARNED_FROZEN = True
                import warnings
                warnings.warn('ood_gan_inversion_b200: generator parameters require grad, but this path treats the generator as frozen '
                              '(options/train/E4E_Face.yml:123-125) -- the output carries no gradient to them; call '
                              'requires_grad_(False) on the generator, or run under torch.no_grad(), to silence this', stacklevel=3)
        if grad_route:
            if kwargs.get('features_in', None) is not None:
                raise NotImplementedError('ood_gan_inversion_b200: the differentiable path does not take features_in (Feature-Style encoder hook)')
            b = latent.shape[0]
            nz = []
            for li in range(self.num_layers):
                res = 2 ** ((li + 5) // 2)
                n = noise[li]
                if has_cond and li in cond_layers:
                    n = conditions[cond_layers.index(li)][1]            # a conditioned up-conv (layer index == latent index): its noise
                                                                        # is conditions[ci][1], drawn when None (model.py:558-571, 277-292)
                nz.append(n.detach().float().contiguous() if n is not None else
                          torch.empty(b, 1, res, res, device=latent.device, dtype=torch.float32).normal_())
            if not has_cond and not return_features:
                # optimisation-based inversion of the plain synthesis: ONE Function, hand-written backward (synthesis_grad)
                from .synthesis_grad import synthesis
                return synthesis(self, latent, nz), None
            # something to differentiate through between the layers (alignment callback, returned features): layer-wise graph
            from . import synthesis_diff
            hooks = {}
            if has_cond:
                lat_f = latent.float()
                for ci, i in enumerate(cond_layers):
                    if conditions[ci][1] is not None or not kwargs.get('callback', None):
                        continue
                    j = i                                        # latent index i = 1 + 2*blk is also the layer index of that up-conv
                    kw = dict(kwargs)
                    kw.update({'index': ci, 'style': lat_f[:, i]})
                    conv1 = self.convs[j - 1]
                    if not hasattr(kw['callback'], 'aligned_nhwc'):
                        raise NotImplementedError('ood_gan_inversion_b200: the differentiable path needs a callback that exposes aligned_nhwc '
                                                  '(this package\'s arch); a foreign NCHW callback would cut the graph')
                    hook, _ = _callback_hook(conv1.noise, nz[j], kw, conv1.conv.out_channel)
                    hooks[j] = hook
            image, y = synthesis_diff.synthesis(self, latent, nz, hooks, return_features)
            feat = _to_nchw(y, self.convs[-1].conv.out_channel if self.log_size > 2 else self.conv1.conv.out_channel) if return_features else None
            return image, feat
        lat = latent.detach().float().contiguous()
        b = lat.shape[0]
        dev = lat.device
        dt = _act_dtype()
        features_in = kwargs.get('features_in', None)
        feature_scale = kwargs.get('feature_scale', 1.0)

        def draw(layer_noise, res):
            if layer_noise is not None:
                return layer_noise.detach().float().contiguous()
            return torch.empty(b, 1, res, res, device=dev, dtype=torch.float32).normal_()

        def insert_feature(y, s_conv, layer_idx):
            # Feature-Style-encoder hook (model.py:541-546): blend in NCHW, then re-enter the pipeline
            if features_in is None or features_in[layer_idx] is None:
                return None
            f = features_in[layer_idx].float()
            x = (1 - feature_scale) * _to_nchw(y, f.shape[1]) + feature_scale * f
            return _to_nhwc(x, s_conv, s_conv.shape[1])

        n_blocks = self.log_size - 2
        # every 3x3 layer's (s, d) up front: all styles are known before the first convolution
        sd = [self.conv1.conv.coeffs(lat[:, 0])] + [c.conv.coeffs(lat[:, 1 + j]) for j, c in enumerate(self.convs)]
        s0, d0 = sd[0]
        xs = _to_nhwc(self.input.input.detach(), s0, self.conv1.conv.cin_p, batch=b)
        s_next = sd[1][0] if n_blocks > 0 else None
        y, ys = self.conv1.run_nhwc(xs, lat[:, 0], draw(noise[0], 4), s_next=s_next, want_y=True, want_ys=n_blocks > 0, d=d0)
        skip = self.to_rgb1.run_nhwc(y, lat[:, 1])
        i = 1
        for blk in range(n_blocks):
            conv1, conv2, to_rgb = self.convs[2 * blk], self.convs[2 * blk + 1], self.to_rgbs[blk]
            res = 2 ** (blk + 3)
            ins = insert_feature(y, s_next, i)
            if ins is not None:
                ys = ins
            d1, (s2, d2) = sd[1 + 2 * blk][1], sd[2 + 2 * blk]
            need_y1 = features_in is not None and features_in[i + 1] is not None
            hook, n1 = None, noise[1 + 2 * blk]
            if cond_layers is not None and conditions is not None and i in cond_layers:
                ci = cond_layers.index(i)
                n1 = conditions[ci][1]
                kw = dict(kwargs)
                kw.update({'index': ci, 'style': lat[:, i]})
                if n1 is None and kw.get('callback', None):
                    n1 = draw(None, res)
                    hook, n1 = _callback_hook(conv1.noise, n1, kw, conv1.conv.out_channel)
                    y1, y1s = conv1.run_nhwc(ys, lat[:, i], n1, s_next=s2, want_y=need_y1, want_ys=True, hook=hook, d=d1)
                else:
                    y1, y1s = conv1.run_nhwc(ys, lat[:, i], draw(n1, res), s_next=s2, want_y=need_y1, want_ys=True, d=d1)
            else:
                y1, y1s = conv1.run_nhwc(ys, lat[:, i], draw(n1, res), s_next=s2, want_y=need_y1, want_ys=True, d=d1)
            if need_y1:
                y1s = insert_feature(y1, s2, i + 1)
            last = blk == n_blocks - 1
            s_next = None if last else sd[3 + 2 * blk][0]
            # ToRGB fused into conv2's epilogue when one N tile holds all channels (tcgen05 path, Co <= 256): the unscaled
            # activation is then neither written nor re-read (it is only kept when someone else needs it)
            co = conv2.conv.cout_p
            need_y = return_features and last or (features_in is not None and not last and features_in[i + 2] is not None)
            # 128/256 channels: the generic tiles' epilogue.  32/64 channels at 512 / 1024 px (the row-sliding kernel): NOT fused by default -- that
            # kernel is bound by its epilogue, and the per-thread dot products + skip taps cost it more than the separate ToRGB pass costs:
            # round 1 measured 0.9 -> 1.8 ms with the eight-warp epilogue; round 2 with the one-pixel-per-thread epilogue: conv time 14.0 -> 15.0 ms
            # for 0.43 ms of ToRGB removed, 639.5 -> 625.6 images/s on the same box (OOD_FUSE_RGB_ROWS=1 repeats the experiment; =2 fuses only the last
            # layer, whose activation is then not written at all: conv +0.85 ms for 0.30 ms of ToRGB, 676-689 -> 661-669 images/s; ncu source page of
            # that kernel (RGB=1 scripts/rows_bench.py: 1293 us against 474 + 290 us separately): 29 % of the stall samples sit on the first use of the
            # up-sampled skip taps, requesting them a row ahead did not move it -- the compiler keeps the dependent FMAs next to the loads)
            rows_ok = (_FUSE_RGB_ROWS == 1 or _FUSE_RGB_ROWS == 2 and last and not need_y) and co in (32, 64) and conv2.conv.cin_p == co and res % 128 == 0 and \
                b * ((res + 31) // 32) * (res // 128) >= int(os.environ.get('OOD_ROWS_MIN_STRIPS', 148))
            fuse = _PRECISION == 'bf16' and (128 <= co <= 256 and _FUSE_RGB_TC or rows_ok) and co == conv2.conv.out_channel and len(to_rgb.taps_up) == 4 and res % 2 == 0
            if fuse:
                y, ys, skip = conv2.run_nhwc(y1s, lat[:, i + 1], draw(noise[2 + 2 * blk], res), s_next=s_next, want_y=need_y,
                                             want_ys=not last, d=d2, rgb=to_rgb.fused_args(lat[:, i + 2], skip))
            else:
                y, ys = conv2.run_nhwc(y1s, lat[:, i + 1], draw(noise[2 + 2 * blk], res), s_next=s_next, want_y=True,
                                       want_ys=not last, d=d2)
                skip = to_rgb.run_nhwc(y, lat[:, i + 2], skip)
            i += 2
        feat = _to_nchw(y, self.convs[-1].conv.out_channel if n_blocks else self.conv1.conv.out_channel) if return_features else None
        return skip, feat
