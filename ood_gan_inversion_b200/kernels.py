"""Tensor-level wrappers over the C ABI (include/ood_b200.h).  PyTorch is only the allocator / stream provider here.

All functions require CUDA tensors; NHWC activations are plain contiguous [B,H,W,C] tensors (fp32 or bf16).
"""
import ctypes as C
import math

import torch

from . import _lib
from ._lib import BF16, F16, F32, BlurActArgs, ConvArgs, check

FIR_1331 = (1.0, 3.0, 3.0, 1.0)

# ---- optional per-launch timing (bench.py roofline): CUDA events on the launching stream ------------------------
_prof = None
_prof_only = None


def profile_begin(only=None):
    """Start recording per-launch CUDA events; `only` restricts it to the named kernels (cheap enough for a timed region)."""
    global _prof, _prof_only
    _prof = []
    _prof_only = set(only) if only else None


def profile_end():
    """-> {name: dict(ms=total device ms, work=total algorithmic flops or bytes, launches=n)}; call after a sync."""
    global _prof
    rec, _prof = _prof, None
    out = {}
    for name, e0, e1, work in rec or []:
        d = out.setdefault(name, dict(ms=0.0, work=0.0, launches=0))
        d['ms'] += e0.elapsed_time(e1)
        d['work'] += work
        d['launches'] += 1
    return out


class _timed:
    def __init__(self, name, work):
        self.name, self.work = name, work

    def __enter__(self):
        self.on = _prof is not None and (_prof_only is None or self.name in _prof_only)
        if self.on:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.on and _prof is not None:
            self.e1.record()
            _prof.append((self.name, self.e0, self.e1, self.work))
        return False


def _esize(t):
    return t.element_size()


def _dt(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float16:      # the encoder path (OOD_F16): conv3x3 impl 0, in_stats, se_residual, bicubic_up_add, thumbnail_nhwc
        return F16
    raise RuntimeError(f'ood_gan_inversion_b200: unsupported dtype {t.dtype} (float32, bfloat16; float16 on the encoder kernels)')


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    """The current torch stream of the CURRENT device: every public wrapper runs under _device_guard, which makes the
    device of its tensor arguments current first (the reference's ops do the same: cudaGetDevice + getCurrentCUDAStream,
    upfirdn2d_kernel.cu:143-145)."""
    return torch.cuda.current_stream().cuda_stream


def _cuda(*ts):
    """Every tensor argument is a CUDA tensor, all on ONE device, and that device is current (see _device_guard)."""
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError('ood_gan_inversion_b200 is CUDA-only: got a CPU tensor (there is no CPU fallback)')
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f'ood_gan_inversion_b200: tensor arguments on different devices ({dev} and {t.device})')
    if dev is not None and dev.index != torch.cuda.current_device():
        raise RuntimeError(f'ood_gan_inversion_b200: tensors on {dev} but the current device is cuda:{torch.cuda.current_device()} '
                           '(internal: a wrapper ran outside _device_guard)')


def _first_cuda_device(args, kwargs):
    for a in list(args) + list(kwargs.values()):
        if isinstance(a, torch.Tensor):
            if a.is_cuda:
                return a.device
        elif isinstance(a, (list, tuple)):
            for e in a:
                if isinstance(e, torch.Tensor) and e.is_cuda:
                    return e.device
    return None


def _device_guard(fn):
    """Run `fn` with the device of its first CUDA tensor argument current, so that launches, the stream handed to the C ABI
    and the tensors' memory agree when a process drives several GPUs (a model on cuda:1 while cuda:0 is current)."""
    import functools

    @functools.wraps(fn)
    def guarded(*args, **kwargs):
        dev = _first_cuda_device(args, kwargs)
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return guarded


def _noise_bstride(noise, b, oh, ow):
    """Batch stride of an injected-noise plane [B|1, 1, oh, ow] fp32 (model.py:277-283 broadcasts a single plane)."""
    if noise is None:
        return 0
    if noise.dtype != torch.float32 or not noise.is_contiguous():
        raise RuntimeError('ood_gan_inversion_b200: noise must be a contiguous float32 tensor')
    if noise.shape[0] not in (1, b) or tuple(noise.shape[-2:]) != (oh, ow) or noise.numel() != noise.shape[0] * oh * ow:
        raise RuntimeError(f'ood_gan_inversion_b200: noise of shape {tuple(noise.shape)} does not fit a batch of {b} planes of {oh}x{ow} '
                           '(expected [B,1,H,W] or [1,1,H,W])')
    return 0 if noise.shape[0] == 1 else oh * ow


def _f32c(t):
    return None if t is None else t.detach().to(torch.float32).contiguous()


def fir_taps(taps=FIR_1331, gain=1.0):
    """1-D FIR taps normalised to sum `gain` (2-D kernel = outer product; model.py:19-27)."""
    s = float(sum(taps))
    return [float(t) / s * gain for t in taps]


# ----------------------------------------------------------------------------------------------- boundary #2
def upfirdn2d_nchw(x, kernel, up_x, up_y, down_x, down_y, px0, px1, py0, py1):
    _cuda(x, kernel)
    b, c, h, w = x.shape
    kh, kw = kernel.shape
    x = x.contiguous()
    k = _f32c(kernel)
    oh = (h * up_y + py0 + py1 - kh) // down_y + 1
    ow = (w * up_x + px0 + px1 - kw) // down_x + 1
    if oh <= 0 or ow <= 0:
        raise RuntimeError(f'upfirdn2d: empty output {oh}x{ow}')
    out = torch.empty(b, c, oh, ow, device=x.device, dtype=x.dtype)
    check(_lib.lib().ood_upfirdn2d(_ptr(x), _ptr(out), _ptr(k), b * c, h, w, kh, kw, up_x, up_y, down_x, down_y,
                                   px0, px1, py0, py1, _dt(x), _stream()), 'upfirdn2d')
    return out


def fused_bias_act(x, bias, refer, grad, alpha, scale):
    """x: [B,C,...] contiguous; bias fp32 [C] or None; refer like x or None."""
    _cuda(x, bias, refer)
    x = x.contiguous()
    out = torch.empty_like(x)
    channels = x.shape[1] if x.dim() > 1 else 1
    inner = 1
    for s in x.shape[2:]:
        inner *= s
    bias = _f32c(bias)
    refer = None if refer is None else refer.contiguous()
    check(_lib.lib().ood_fused_bias_act(_ptr(x), _ptr(bias), _ptr(refer), _ptr(out), x.numel(), channels, inner, 3, grad,
                                        float(alpha), float(scale), _dt(x), _stream()), 'fused_bias_act')
    return out


def bias_grad(g):
    _cuda(g)
    g = g.contiguous()
    channels = g.shape[1]
    inner = 1
    for s in g.shape[2:]:
        inner *= s
    out = torch.empty(channels, device=g.device, dtype=torch.float32)
    check(_lib.lib().ood_bias_grad(_ptr(g), _ptr(out), g.shape[0], channels, inner, _dt(g), _stream()), 'bias_grad')
    return out


# ----------------------------------------------------------------------------------------------- layout
def nchw_to_nhwc(x, scale_bc=None, dtype=torch.bfloat16, batch=None):
    """x: fp32 [B,C,H,W] (or [1,C,H,W] broadcast to `batch`) -> [B,H,W,C] `dtype`, times scale_bc[b,c]."""
    _cuda(x, scale_bc)
    x = _f32c(x)
    b0, c, h, w = x.shape
    b = batch or b0
    bstride = 0 if (b0 == 1 and b > 1) else c * h * w
    out = torch.empty(b, h, w, c, device=x.device, dtype=dtype)
    check(_lib.lib().ood_nchw_to_nhwc(_ptr(x), bstride, _ptr(_f32c(scale_bc)), _ptr(out), b, c, h, w, _dt(out), _stream()),
          'nchw_to_nhwc')
    return out


def thumbnail_nhwc(x, size=(256, 256), cp=32, dtype=torch.bfloat16):
    """x fp32 NCHW [B,3,H,W] -> `dtype` NHWC [B,oh,ow,cp]: F.interpolate(x, size, mode='bilinear') (e4e_arch.py:256) written as the
    encoder's first-convolution operand (channels 3..cp-1 zero)."""
    _cuda(x)
    x = _f32c(x)
    b, c, h, w = x.shape
    oh, ow = size
    out = torch.empty(b, oh, ow, cp, device=x.device, dtype=dtype)
    with _timed('thumbnail', b * (4 * c * oh * ow * 4 + oh * ow * cp * out.element_size())):
        check(_lib.lib().ood_thumbnail_nhwc(_ptr(x), _ptr(out), b, c, h, w, oh, ow, cp, _dt(out), _stream()), 'thumbnail_nhwc')
    return out


def nhwc_to_nchw(x):
    _cuda(x)
    x = x.contiguous()
    b, h, w, c = x.shape
    out = torch.empty(b, c, h, w, device=x.device, dtype=torch.float32)
    check(_lib.lib().ood_nhwc_to_nchw(_ptr(x), _ptr(out), b, c, h, w, _dt(x), _stream()), 'nhwc_to_nchw')
    return out


def nhwc_scale(x, scale_bc):
    _cuda(x, scale_bc)
    x = x.contiguous()
    b, h, w, c = x.shape
    out = torch.empty_like(x)
    check(_lib.lib().ood_nhwc_scale(_ptr(x), _ptr(_f32c(scale_bc)), _ptr(out), b, c, h * w, _dt(x), _stream()), 'nhwc_scale')
    return out


# ----------------------------------------------------------------------------------------------- modulation
def weight_sumsq(w):
    """w fp32 [Co,Ci,kh,kw] -> [Co,Ci]"""
    _cuda(w)
    w = _f32c(w)
    co, ci = w.shape[:2]
    out = torch.empty(co, ci, device=w.device, dtype=torch.float32)
    check(_lib.lib().ood_weight_sumsq(_ptr(w), _ptr(out), co, ci, w.shape[2] * w.shape[3], _stream()), 'weight_sumsq')
    return out


def pack_conv_weight(w, dtype, ci_major):
    """w fp32 [Co,Ci,kh,kw] -> [taps,Co,Ci] (`dtype`) or [taps,Ci,Co] fp32 (ci_major)."""
    _cuda(w)
    w = _f32c(w)
    co, ci = w.shape[:2]
    taps = w.shape[2] * w.shape[3]
    shape = (taps, ci, co) if ci_major else (taps, co, ci)
    out = torch.empty(shape, device=w.device, dtype=dtype)
    check(_lib.lib().ood_pack_conv_weight(_ptr(w), _ptr(out), co, ci, taps, int(ci_major), _dt(out), _stream()),
          'pack_conv_weight')
    return out


def pack_convt_fused(w9):
    """[9][Co][Ci] bf16 pack (pack_conv_weight) -> [4 shifts][4*Co][Ci] pack of the fused-phase transposed form (conv3x3
    transposed=5, see ood_b200.h): shift t reads input (oy - (t>>1), ox - (t&1)); row block 2*py + px is output parity (py, px)."""
    _, co, ci = w9.shape
    out = torch.zeros(4, 4 * co, ci, device=w9.device, dtype=w9.dtype)
    for t in range(4):
        sy, sx = t >> 1, t & 1                      # 1 = the shift by -1
        for py in range(2):
            for px in range(2):
                ky = py if sy == 0 else (2 if py == 0 else None)
                kx = px if sx == 0 else (2 if px == 0 else None)
                if ky is None or kx is None:
                    continue
                ph = 2 * py + px
                out[t, ph * co:(ph + 1) * co] = w9[ky * 3 + kx]
    return out.contiguous()


def modulation(latent, mod_w, mod_b, wsq, conv_scale, cout, want_d=True):
    """latent [B,D] fp32 (row stride may exceed D) -> s [B,Ci], d [B,Co] (or None)."""
    _cuda(latent, mod_w, mod_b, wsq)
    assert latent.dim() == 2 and latent.stride(1) == 1 and latent.dtype == torch.float32
    b, dim = latent.shape
    cin = mod_w.shape[0]
    s = torch.empty(b, cin, device=latent.device, dtype=torch.float32)
    d = torch.empty(b, cout, device=latent.device, dtype=torch.float32) if want_d else None
    check(_lib.lib().ood_modulation(_ptr(latent), latent.stride(0), _ptr(mod_w), _ptr(mod_b), _ptr(wsq), float(conv_scale),
                                    _ptr(s), _ptr(d), b, dim, cin, cout, _stream()), 'modulation')
    return s, d


def torgb_weight(w, s, scale=None):
    """w fp32 [3,C], s [B,C] -> [B,3,C] = w*s*scale (scale defaults to 1/sqrt(C))"""
    _cuda(w, s)
    b, c = s.shape
    out = torch.empty(b, 3, c, device=s.device, dtype=torch.float32)
    scale = 1.0 / math.sqrt(c) if scale is None else float(scale)
    check(_lib.lib().ood_torgb_weight(_ptr(w), _ptr(s), _ptr(out), scale, b, c, _stream()), 'torgb_weight')
    return out


# ----------------------------------------------------------------------------------------------- convolution
def conv3x3(x, weight, cout, transposed=False, impl=0, d=None, noise=None, noise_w=None, bias=None, s_next=None,
            act=False, want_y=True, want_ys=False, out_f32=False, prelu=None, rgb=None, tag=None, groups=1, in_shared=False,
            acc_in=None, tiled=False, stats_eps=None, out_dtype=None, out=None, tile_sums=False):
    """x NHWC [B,H,W,Ci] pre-modulated; weight packed for `impl`.  Returns (y, ys) (None where not requested).
    rgb=(wrgb [B,3,Co], bias [3], skip NCHW fp32 or None, up taps): fused ToRGB, returns (y, ys, rgb_out NCHW fp32).
    acc_in: fp32 NHWC [B,OH,OW,Co] seed added to the accumulator before the epilogue (tcgen05 path).
    tiled: acc_in, and y when out_f32, are flat fp32 tensors in the kernel's tile order (ood_conv3x3_tiled_bytes): the fast form
    of a seed that only ever travels between two launches with the same geometry.
    stats_eps: also return [B,Co,2] = (mean, rstd) of y as stored, computed in the epilogue -> (y, ys, stats); use
    conv3x3_stats_ok() for the envelope.  tile_sums=True: (y, ys, ws) with ws = the epilogue's per-tile channel sums [B, tiles, Co, 2]
    (sum of y as stored, 0), no finalise launch: the pooled sums for se_apply().
    out_dtype: storage type of y / ys when it differs from x's (torch.bfloat16 or torch.float16; tcgen05 path).
    out: a preallocated contiguous tensor to receive y (shape / dtype checked)."""
    _cuda(x, weight, d, noise, noise_w, bias, s_next, acc_in)
    assert x.is_contiguous()
    if weight.dtype != x.dtype and impl == 0:
        raise RuntimeError(f'ood_gan_inversion_b200: conv3x3 operands differ in type ({x.dtype} activations, {weight.dtype} weights)')
    odt = x.dtype if out_dtype is None else out_dtype
    if odt != x.dtype and (impl != 0 or odt not in (torch.bfloat16, torch.float16) or x.dtype == torch.float32):
        raise RuntimeError('ood_gan_inversion_b200: conv3x3 out_dtype is a bf16 <-> f16 switch of the tcgen05 path')
    b, h, w, cin = x.shape
    if in_shared:            # grouped form on a shared input: every group convolves the same [images, ...] tensor
        b = b * groups
    transposed = int(transposed)      # 0 stride-1 | 1 stride-2 transposed | 2 stride-2 valid (data gradient of 1) | 3 stride-2 pad-1 | 4 1x1 | 5 fused 1 | 6 1x1 stride-2
    oh, ow = {0: (h, w), 1: (2 * h + 1, 2 * w + 1), 2: ((h - 1) // 2, (w - 1) // 2),
              3: ((h - 1) // 2 + 1, (w - 1) // 2 + 1), 4: (h, w), 5: (2 * h + 1, 2 * w + 1),
              6: ((h - 1) // 2 + 1, (w - 1) // 2 + 1)}[transposed]
    nt = 0
    if tiled:
        nt = _lib.lib().ood_conv3x3_tiled_bytes(b, h, w, cin, cout, transposed) // 4
        if nt <= 0:
            raise RuntimeError('ood_gan_inversion_b200: conv3x3 tiled=True needs cout % 128 == 0 and a single-phase form')
    if tiled and out_f32 and want_y:
        y = torch.empty(nt, device=x.device, dtype=torch.float32)
    else:
        y = torch.empty(b, oh, ow, cout, device=x.device, dtype=torch.float32 if out_f32 else odt) if want_y else None
    if out is not None:
        _cuda(out)
        if y is None or out.shape != y.shape or out.dtype != y.dtype or not out.is_contiguous():
            raise RuntimeError(f'ood_gan_inversion_b200: conv3x3 out= must be a contiguous {None if y is None else (tuple(y.shape), y.dtype)} tensor')
        y = out
    ys = torch.empty(b, oh, ow, cout, device=x.device, dtype=odt) if want_ys else None
    nbs = _noise_bstride(noise, b, oh, ow)
    a = ConvArgs(_ptr(x), _ptr(weight), _ptr(y), _ptr(ys), _ptr(d), _ptr(noise), nbs, _ptr(noise_w), _ptr(bias), _ptr(s_next),
                 b, h, w, cin, cout, int(transposed), 2 if prelu is not None else int(act), impl, _dt(x), int(out_f32), _ptr(prelu))
    a.groups, a.in_shared = int(groups), int(bool(in_shared))
    a.out_dtype = 0 if odt == x.dtype else (F16 if odt == torch.float16 else BF16)
    if acc_in is not None:
        assert acc_in.dtype == torch.float32 and acc_in.is_contiguous()
        assert tuple(acc_in.shape) == ((nt,) if tiled else (b, oh, ow, cout))
        a.acc_in = _ptr(acc_in)
    a.tiled = int(bool(tiled))
    st = None
    if stats_eps is not None:
        nws = _lib.lib().ood_conv3x3_stats_workspace(b, h, w, cin, cout, transposed) // 4
        if nws <= 0:
            raise RuntimeError('ood_gan_inversion_b200: conv3x3 fused statistics are outside their envelope (see ood_b200.h)')
        ws = torch.empty(nws, device=x.device, dtype=torch.float32)
        st = torch.empty(b, cout, 2, device=x.device, dtype=torch.float32)
        a.stats_ws, a.stats_out, a.stats_eps = _ptr(ws), _ptr(st), float(stats_eps)
    if tile_sums:
        assert stats_eps is None
        nws = _lib.lib().ood_conv3x3_stats_workspace(b, h, w, cin, cout, transposed) // 4
        if nws <= 0:
            raise RuntimeError('ood_gan_inversion_b200: conv3x3 fused statistics are outside their envelope (see ood_b200.h)')
        st = torch.empty(b, nws // (b * cout * 2), cout, 2, device=x.device, dtype=torch.float32)
        a.stats_ws = _ptr(st)
    rgb_out = None
    if rgb is not None:
        wrgb, rbias, rskip, rtaps = rgb
        _cuda(wrgb, rbias, rskip)
        rgb_out = torch.empty(b, 3, oh, ow, device=x.device, dtype=torch.float32)
        a.rgb_w, a.rgb_bias, a.rgb_skip, a.rgb_out = _ptr(wrgb), _ptr(rbias), _ptr(rskip), _ptr(rgb_out)
        a.rgb_taps = (C.c_float * 4)(*rtaps)
    # algorithmic work: 2*B*Co*Ci*9*H*W with H, W the INPUT size for the transposed form
    px = h * w if transposed in (0, 1, 5) else oh * ow
    with _timed(tag or ('conv3x3_tc' if impl == 0 else 'conv3x3_simt'), 2.0 * b * cout * cin * (1 if transposed in (4, 6) else 9) * px) as tm:
        check(_lib.lib().ood_conv3x3(C.byref(a), _stream()), 'conv3x3')
        if tm.on and tag is None and _lib.lib().ood_last_conv_route() in (1, 2):
            # the HBM-bound row kernels (conv_rows.cu / convt_rows.cu) are accounted in bytes under their own name: input + every output tensor + noise
            out_bytes = sum(t.numel() * t.element_size() for t in (y, ys) if t is not None)
            tm.name = 'conv3x3_rows'
            tm.work = float(b * h * w * cin * _esize(x) + out_bytes + (b * oh * ow * 4 if noise is not None else 0))
    if rgb is not None:
        return y, ys, rgb_out
    if stats_eps is not None or tile_sums:
        return y, ys, st
    return y, ys


def conv3x3_stats_ok(x, cout, transposed=0):
    """True when conv3x3(..., stats_eps=... / tile_sums=True) is available for this input (tcgen05 path, 16-bit storage, wide tiles, one
    image per tile)."""
    b, h, w, cin = x.shape
    return x.dtype in (torch.bfloat16, torch.float16) and cin % 64 == 0 and \
        _lib.lib().ood_conv3x3_stats_workspace(b, h, w, cin, cout, int(transposed)) > 0


def blur_act(t, taps, d=None, noise=None, noise_w=None, bias=None, s_next=None, act=True, want_img=False, want_y=True,
             want_ys=False, dtype=None, pad=(1, 1)):
    """t NHWC [B,IH,IW,C] (fp32 or storage dtype) -> blurred [B,IH+p0+p1-3,...,C] (+ fused StyledConv tail).
    pad (1,1): the blur after the transposed conv; pad (2,2): its adjoint (gradient w.r.t. the transposed-conv output)."""
    _cuda(t, d, noise, noise_w, bias, s_next)
    assert t.is_contiguous()
    b, ih, iw, c = t.shape
    dtype = dtype or t.dtype
    oh, ow = ih + pad[0] + pad[1] - 3, iw + pad[0] + pad[1] - 3
    mk = lambda: torch.empty(b, oh, ow, c, device=t.device, dtype=dtype)
    img = mk() if want_img else None
    y = mk() if (act and want_y) else None
    ys = mk() if (act and want_ys) else None
    nbs = _noise_bstride(noise, b, oh, ow)
    a = BlurActArgs(_ptr(t), int(t.dtype == torch.float32 and dtype != torch.float32), _ptr(img), _ptr(y), _ptr(ys), _ptr(d),
                    _ptr(noise), _ptr(noise_w), _ptr(bias), _ptr(s_next), nbs, (C.c_float * 4)(*taps), b, ih, iw, c,
                    int(act), F32 if dtype == torch.float32 else BF16, int(pad[0]), int(pad[1]))
    nout = sum(o is not None for o in (img, y, ys))
    work = b * c * (ih * iw * _esize(t) + nout * oh * ow * (4 if dtype == torch.float32 else 2)) + \
        (0 if noise is None else noise.shape[0] * oh * ow * 4)
    with _timed('blur_act', work):
        check(_lib.lib().ood_blur_act(C.byref(a), _stream()), 'blur_act')
    return img, y, ys


def noise_act(img, noise, noise_w, bias, s_next=None, want_y=True, want_ys=False):
    _cuda(img, noise, noise_w, bias, s_next)
    assert img.is_contiguous()
    b, h, w, c = img.shape
    y = torch.empty_like(img) if want_y else None
    ys = torch.empty_like(img) if want_ys else None
    nbs = _noise_bstride(noise, b, h, w)
    nout = (y is not None) + (ys is not None)
    with _timed('noise_act', b * h * w * (c * _esize(img) * (1 + nout) + 4)):
        check(_lib.lib().ood_noise_act(_ptr(img), _ptr(y), _ptr(ys), _ptr(noise), nbs, _ptr(noise_w), _ptr(bias), _ptr(s_next),
                                       b, h * w, c, _dt(img), _stream()), 'noise_act')
    return y, ys


def torgb(y, wrgb, bias, skip=None, taps_up=None):
    """y NHWC [B,H,W,C]; wrgb [B,3,C]; bias [3]; skip NCHW fp32 [B,3,H/2,W/2] -> NCHW fp32 [B,3,H,W]"""
    _cuda(y, wrgb, bias, skip)
    assert y.is_contiguous()
    b, h, w, c = y.shape
    out = torch.empty(b, 3, h, w, device=y.device, dtype=torch.float32)
    taps = (C.c_float * 4)(*(taps_up or fir_taps(gain=2.0)))
    work = b * h * w * (c * _esize(y) + 3 * 4 + (0 if skip is None else 3)) 
    with _timed('torgb', work):
        check(_lib.lib().ood_torgb(_ptr(y), _ptr(wrgb), _ptr(_f32c(bias).reshape(-1)), _ptr(skip), _ptr(out), taps, b, h, w, c,
                                   _dt(y), _stream()), 'torgb')
    return out


# ----------------------------------------------------------------------------------------------- SAMM
def field_step(z, prev, coarse, scale, taps=None, z2=None, coef=None):
    """z fp32 [B,3,R,R] pre-activation (or z*coef0 + z2*coef1 + coef2 with coef [B,3,3] from alignnet_tail)."""
    _cuda(z, prev, coarse, z2, coef)
    z = _f32c(z)
    b, _, r, _ = z.shape
    acc = torch.empty_like(z)
    rc = coarse.shape[-1] if coarse is not None else 0
    t = (C.c_float * 4)(*(taps or fir_taps()))
    check(_lib.lib().ood_field_step(_ptr(z), _ptr(_f32c(prev)), _ptr(_f32c(coarse)), _ptr(acc), t, float(scale), b, r, rc,
                                    _ptr(_f32c(z2)), _ptr(_f32c(coef)), _stream()), 'field_step')
    return acc


def alignnet_tail(res, shortcut, slope, conv_w, in_res_w, in_res_b, in_sc_w, in_sc_b, eps=1e-5):
    """-> (r2, coef): InstanceNorm(conv3x3(PReLU(res))) + InstanceNorm(shortcut) == r2*coef[...,0] + shortcut*coef[...,1] + coef[...,2]"""
    _cuda(res, shortcut, slope, conv_w, in_res_w, in_res_b, in_sc_w, in_sc_b)
    assert res.dtype == torch.float32 and shortcut.dtype == torch.float32 and res.is_contiguous() and shortcut.is_contiguous()
    b, _, r, _ = res.shape
    r2 = torch.empty_like(res)
    ws = torch.empty(_lib.lib().ood_alignnet_tail_workspace(b, r) // 4, device=res.device, dtype=torch.float32)
    coef = torch.empty(b, 3, 3, device=res.device, dtype=torch.float32)
    with _timed('alignnet_tail', b * 3 * r * r * 4 * 3):
        check(_lib.lib().ood_alignnet_tail(_ptr(res), _ptr(shortcut), _ptr(slope), _ptr(conv_w), _ptr(in_res_w), _ptr(in_res_b),
                                           _ptr(in_sc_w), _ptr(in_sc_b), float(eps), _ptr(r2), _ptr(ws), _ptr(coef), b, r, _stream()),
              'alignnet_tail')
    return r2, coef


def warp_mix(gen, field):
    """gen NHWC [B,H,W,C]; field fp32 [B,3,H,W] -> NHWC"""
    _cuda(gen, field)
    assert gen.is_contiguous()
    b, h, w, c = gen.shape
    out = torch.empty_like(gen)
    with _timed('warp_mix', b * h * w * (2 * c * _esize(gen) + 3 * 4)):
        check(_lib.lib().ood_warp_mix(_ptr(gen), _ptr(_f32c(field)), _ptr(out), b, h, w, c, _dt(gen), _stream()), 'warp_mix')
    return out


def _bwd_ws(b, pixels, c, k, device):
    return torch.empty(_lib.lib().ood_bwd_workspace(b, pixels, c, k) // 4, device=device, dtype=torch.float32)


def act_bwd(gy, y, d, bias, noise, noise_w):
    """Backward of lrelu*sqrt2 / bias / noise / demod.  gy, y NHWC -> (g = gv*d NHWC, gd [B,C] fp32)."""
    _cuda(gy, y, d, bias, noise, noise_w)
    assert gy.is_contiguous() and y.is_contiguous() and gy.dtype == y.dtype
    b, h, w, c = y.shape
    g = torch.empty_like(y)
    gd = torch.empty(b, c, device=y.device, dtype=torch.float32)
    nbs = _noise_bstride(noise, b, h, w)
    ws = _bwd_ws(b, h * w, c, 1, y.device)          # kept alive until the launch has been queued
    with _timed('act_bwd', b * h * w * c * _esize(y) * 3):
        check(_lib.lib().ood_act_bwd(_ptr(gy), _ptr(y), _ptr(d), _ptr(bias), _ptr(noise), nbs, _ptr(noise_w), _ptr(g),
                                     _ptr(ws), _ptr(gd), b, h * w, c, _dt(y), _stream()), 'act_bwd')
    return g, gd


def act_bwd_fused(g_in, g_scale, rgb, y, d, bias, noise, noise_w):
    """One pass per layer of the latent-gradient chain (ood_act_bwd_fused): gy = g_in * g_scale + ToRGB gradient, then act_bwd.
    g_in NHWC (unscaled data gradient of the layer above) or None; g_scale [B,C] or None; rgb = (g_rgb [B,3,H,W] fp32, wrgb [B,3,C]) or None.
    Returns (g NHWC, gd [B,C], dot [B,C] = sum_pix g_in*y, g_wrgb [B,3,C] or None)."""
    g_rgb, wrgb = rgb if rgb is not None else (None, None)
    _cuda(g_in, g_scale, g_rgb, wrgb, y, d, bias, noise, noise_w)
    assert y.is_contiguous() and (g_in is None or (g_in.is_contiguous() and g_in.dtype == y.dtype and g_in.shape == y.shape))
    assert g_in is not None or g_rgb is not None
    b, h, w, c = y.shape
    g_rgb, wrgb, g_scale = _f32c(g_rgb), _f32c(wrgb), _f32c(g_scale)
    if g_rgb is not None:
        assert g_rgb.shape == (b, 3, h, w) and wrgb.shape == (b, 3, c)
    k = 5 if g_rgb is not None else 2
    g = torch.empty_like(y)
    sums = torch.empty(b, c, k, device=y.device, dtype=torch.float32)
    nbs = _noise_bstride(noise, b, h, w)
    ws = _bwd_ws(b, h * w, c, k, y.device)
    with _timed('act_bwd_fused', b * h * w * (c * _esize(y) * (3 if g_in is not None else 2) + (12 if g_rgb is not None else 0))):
        check(_lib.lib().ood_act_bwd_fused(_ptr(g_in), _ptr(g_scale), _ptr(g_rgb), _ptr(wrgb), _ptr(y), _ptr(d), _ptr(bias), _ptr(noise), nbs,
                                           _ptr(noise_w), _ptr(g), _ptr(ws), _ptr(sums), b, h * w, c, _dt(y), _stream()), 'act_bwd_fused')
    return g, sums[..., 0], sums[..., 1], (sums[..., 2:5].permute(0, 2, 1) if k == 5 else None)


def dot_reduce(a, x):
    """sum over pixels of a*x per (b, c): NHWC, NHWC -> [B,C] fp32"""
    _cuda(a, x)
    assert a.is_contiguous() and x.is_contiguous() and a.dtype == x.dtype and a.shape == x.shape
    b, h, w, c = a.shape
    out = torch.empty(b, c, device=a.device, dtype=torch.float32)
    ws = _bwd_ws(b, h * w, c, 1, a.device)
    with _timed('dot_reduce', b * h * w * c * _esize(a) * 2):
        check(_lib.lib().ood_dot_reduce(_ptr(a), _ptr(x), _ptr(ws), _ptr(out), b, h * w, c, _dt(a), _stream()), 'dot_reduce')
    return out


def torgb_bwd(g_rgb, wrgb, y, g_in=None):
    """g_rgb NCHW fp32 [B,3,H,W]; wrgb [B,3,C]; y NHWC -> (g_y NHWC = g_in + expand, g_wrgb [B,3,C] fp32)"""
    _cuda(g_rgb, wrgb, y, g_in)
    assert y.is_contiguous() and (g_in is None or (g_in.is_contiguous() and g_in.dtype == y.dtype))
    g_rgb = _f32c(g_rgb)
    b, h, w, c = y.shape
    g_out = torch.empty_like(y)
    gw = torch.empty(b, c, 3, device=y.device, dtype=torch.float32)
    ws = _bwd_ws(b, h * w, c, 3, y.device)
    with _timed('torgb_bwd', b * h * w * (c * _esize(y) * 3 + 12)):
        check(_lib.lib().ood_torgb_bwd(_ptr(g_rgb), _ptr(wrgb), _ptr(y), _ptr(g_in), _ptr(g_out), _ptr(ws), _ptr(gw), b, h * w, c,
                                       _dt(y), _stream()), 'torgb_bwd')
    return g_out, gw.permute(0, 2, 1).contiguous()


def in_stats(x, y=None, eps=1e-5):
    """x (and y) NHWC [B,H,W,C] -> stats [B,C,2] = (mean, rstd), or for a pair [B,C,6] (see ood_b200.h)."""
    _cuda(x, y)
    assert x.is_contiguous() and (y is None or (y.is_contiguous() and y.shape == x.shape and y.dtype == x.dtype))
    b, h, w, c = x.shape
    ws = torch.empty(_lib.lib().ood_in_stats_workspace(b, h * w, c, int(y is not None)) // 4, device=x.device, dtype=torch.float32)
    st = torch.empty(b, c, 6 if y is not None else 2, device=x.device, dtype=torch.float32)
    with _timed('in_stats', b * h * w * c * _esize(x) * (2 if y is not None else 1)):
        check(_lib.lib().ood_in_stats(_ptr(x), _ptr(y), _ptr(ws), _ptr(st), b, h * w, c, float(eps), _dt(x), _stream()), 'in_stats')
    return st


def alignnet_front(cur, enc, st6, w, bias):
    _cuda(cur, enc, st6, w, bias)
    b, h, wd, c = cur.shape
    out = torch.empty(b, h, wd, 2 * c, device=cur.device, dtype=cur.dtype)
    with _timed('alignnet_ew', b * h * wd * c * _esize(cur) * 4):
        check(_lib.lib().ood_alignnet_front(_ptr(cur), _ptr(enc), _ptr(st6), _ptr(w), _ptr(bias), _ptr(out), b, h * wd, c,
                                            _dt(cur), _stream()), 'alignnet_front')
    return out


def alignnet_front_split(cur, enc, st6, w, bias, want_hi=True):
    """The two halves of alignnet_front as separate [B,H,W,C] tensors: (lo, hi); hi (the IN(enc) half, independent of `cur`)
    is skipped with want_hi=False."""
    _cuda(cur, enc, st6, w, bias)
    b, h, wd, c = cur.shape
    lo = torch.empty_like(cur)
    hi = torch.empty_like(cur) if want_hi else None
    with _timed('alignnet_ew', b * h * wd * c * _esize(cur) * (4 if want_hi else 3)):
        check(_lib.lib().ood_alignnet_front_split(_ptr(cur), _ptr(enc), _ptr(st6), _ptr(w), _ptr(bias), _ptr(lo), _ptr(hi), b, h * wd, c,
                                                  _dt(cur), _stream()), 'alignnet_front_split')
    return lo, hi


def alignnet_res0_stats(t, st2, w, bias, cur, enc, st6, eps=1e-5):
    """alignnet_res0 + the (mean, rstd) [B,2C,2] of its output as stored, from the same pass."""
    _cuda(t, st2, w, bias, cur, enc, st6)
    b, h, wd, c = cur.shape
    out = torch.empty_like(t)
    ws = torch.empty(_lib.lib().ood_alignnet_res0_workspace(b, h * wd, c, _dt(cur)) // 4, device=cur.device, dtype=torch.float32)
    st = torch.empty(b, 2 * c, 2, device=cur.device, dtype=torch.float32)
    with _timed('alignnet_ew', b * h * wd * c * _esize(cur) * 6):
        check(_lib.lib().ood_alignnet_res0_stats(_ptr(t), _ptr(st2), _ptr(w), _ptr(bias), _ptr(cur), _ptr(enc), _ptr(st6), _ptr(out),
                                                 _ptr(ws), _ptr(st), float(eps), b, h * wd, c, _dt(cur), _stream()), 'alignnet_res0_stats')
    return out, st


def alignnet_res0(t, st2, w, bias, cur, enc, st6):
    _cuda(t, st2, w, bias, cur, enc, st6)
    b, h, wd, c = cur.shape
    out = torch.empty_like(t)
    with _timed('alignnet_ew', b * h * wd * c * _esize(cur) * 6):
        check(_lib.lib().ood_alignnet_res0(_ptr(t), _ptr(st2), _ptr(w), _ptr(bias), _ptr(cur), _ptr(enc), _ptr(st6), _ptr(out),
                                           b, h * wd, c, _dt(cur), _stream()), 'alignnet_res0')
    return out


def in_apply(x, st2, w=None, bias=None):
    _cuda(x, st2, w, bias)
    b, h, wd, c = x.shape
    out = torch.empty_like(x)
    with _timed('alignnet_ew', b * h * wd * c * _esize(x) * 2):
        check(_lib.lib().ood_in_apply(_ptr(x), _ptr(st2), _ptr(w), _ptr(bias), _ptr(out), b, h * wd, c, _dt(x), _stream()), 'in_apply')
    return out


def bicubic_up_add(x, y):
    """x NHWC [B,h,w,C], y NHWC [B,H,W,C] -> bicubic_up(x, align_corners=True) + y (NHWC)"""
    _cuda(x, y)
    assert x.is_contiguous() and y.is_contiguous() and x.dtype == y.dtype
    b, h, w, c = x.shape
    _, H, W, _ = y.shape
    out = torch.empty_like(y)
    check(_lib.lib().ood_bicubic_up_add(_ptr(x), _ptr(y), _ptr(out), b, h, w, H, W, c, _dt(x), _stream()), 'bicubic_up_add')
    return out


def img2tensor_u8(frames, swap_rb=True, sub=0.5, mul=2.0):
    """frames uint8 [B,H,W,3] (cv2 layout) -> fp32 [B,3,H,W] = (float32(v / 255.0) - sub) * mul, channels reversed if swap_rb."""
    _cuda(frames)
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3 or not frames.is_contiguous():
        raise RuntimeError('img2tensor_u8: expected a contiguous uint8 [B,H,W,3] tensor')
    b, h, w, _ = frames.shape
    out = torch.empty(b, 3, h, w, device=frames.device, dtype=torch.float32)
    with _timed('img2tensor_u8', b * h * w * 15):
        check(_lib.lib().ood_img2tensor_u8(_ptr(frames), _ptr(out), b, h, w, int(bool(swap_rb)), float(sub), float(mul), _stream()),
              'img2tensor_u8')
    return out


def tensor2img_u8(t, swap_rb=True, lo=0.0, hi=1.0):
    """t fp32 [B,3,H,W] -> uint8 [B,H,W,3] = rint(((clamp(t, lo, hi) - lo) / (hi - lo)) * 255), channels reversed if swap_rb."""
    _cuda(t)
    if t.dtype != torch.float32 or t.dim() != 4 or t.shape[1] != 3 or not t.is_contiguous():
        raise RuntimeError('tensor2img_u8: expected a contiguous fp32 [B,3,H,W] tensor')
    b, _, h, w = t.shape
    out = torch.empty(b, h, w, 3, device=t.device, dtype=torch.uint8)
    with _timed('tensor2img_u8', b * h * w * 15):
        check(_lib.lib().ood_tensor2img_u8(_ptr(t), _ptr(out), b, h, w, int(bool(swap_rb)), float(lo), float(hi), _stream()),
              'tensor2img_u8')
    return out


def pack_conv1x1_weight(w, dtype, ci_major):
    """w [Co,Ci] (or [Co,Ci,1,1]) -> the one-tap pack of ood_conv3x3(transposed=4): [1][Co][Ci] (tcgen05) / [1][Ci][Co] (simt)."""
    w = w.reshape(w.shape[0], -1).float()
    return (w.t() if ci_major else w).contiguous().to(dtype).unsqueeze(0)


def tap_sum(proj, shortcut=False):
    """proj fp32 NHWC [B,H,W,Cp>=27] (channel 3*tap + colour) -> fp32 NCHW [B,3,H,W]: the nine shifted partial sums of a 3x3 conv.
    shortcut=True also returns channels 27..29 of every pixel as a second [B,3,H,W] tensor (a 1x1 convolution that rode along)."""
    _cuda(proj)
    assert proj.is_contiguous() and proj.dtype == torch.float32
    b, h, w, cp = proj.shape
    out = torch.empty(b, 3, h, w, device=proj.device, dtype=torch.float32)
    if shortcut:
        sc = torch.empty_like(out)
        with _timed('tap_sum', b * h * w * (30 + 6) * 4):
            check(_lib.lib().ood_tap_sum_shortcut(_ptr(proj), _ptr(out), _ptr(sc), b, h, w, cp, _stream()), 'tap_sum_shortcut')
        return out, sc
    with _timed('tap_sum', b * h * w * (27 + 3) * 4):
        check(_lib.lib().ood_tap_sum(_ptr(proj), _ptr(out), b, h, w, cp, _stream()), 'tap_sum')
    return out


def se_gate(stats, w1, w2):
    """stats [B,C,2] (in_stats) -> gate [B,C] = sigmoid(w2 . relu(w1 . mean))   (SEModule, helpers.py:59-76)"""
    _cuda(stats, w1, w2)
    b, c, _ = stats.shape
    gate = torch.empty(b, c, device=stats.device, dtype=torch.float32)
    check(_lib.lib().ood_se_gate(_ptr(stats), _ptr(w1), _ptr(w2), _ptr(gate), b, c, w1.shape[0], _stream()), 'se_gate')
    return gate


def se_residual(v, gate=None, shortcut=None, sc_stride=1, bn_g=None, bn_h=None, want_out=True, out_f32=False, want_lp=False):
    """v NHWC; out = v*gate + shortcut[:, ::s, ::s]; optionally t_next = out*bn_g + bn_h.  Returns (out, t_next), or with
    want_lp (out, t_next, out_lp): `out` once more in v's storage type (the tapped feature maps of an fp32 residual stream).
    With bf16 / f16 activations the shortcut may be fp32 and `out` can be requested in fp32 (fp32 residual stream)."""
    _cuda(v, gate, shortcut, bn_g, bn_h)
    assert v.is_contiguous() and (shortcut is None or shortcut.is_contiguous())
    b, h, w, c = v.shape
    sc_f32 = shortcut is not None and shortcut.dtype == torch.float32 and v.dtype != torch.float32
    if shortcut is not None:
        assert shortcut.dtype in (v.dtype, torch.float32)
        assert shortcut.shape == (b, h * sc_stride, w * sc_stride, c), 'shortcut must be exactly stride x the output size'
    out_f32 = bool(out_f32) and v.dtype != torch.float32
    out = torch.empty(v.shape, device=v.device, dtype=torch.float32 if out_f32 else v.dtype) if want_out else None
    tn = torch.empty_like(v) if bn_g is not None else None
    lp = torch.empty_like(v) if want_lp else None
    es = _esize(v)
    nbytes = b * h * w * c * (es + (0 if shortcut is None else (4 if sc_f32 else es)) + (0 if out is None else (4 if out_f32 else es)) +
                              (0 if tn is None else es) + (0 if lp is None else es))
    with _timed('se_residual', nbytes):
        check(_lib.lib().ood_se_residual(_ptr(v), _ptr(gate), _ptr(shortcut), int(sc_stride), _ptr(bn_g), _ptr(bn_h), _ptr(out),
                                         _ptr(tn), _ptr(lp), b, h, w, c, _dt(v), int(sc_f32), int(out_f32), _stream()), 'se_residual')
    return (out, tn, lp) if want_lp else (out, tn)


def se_tail(v, w1, w2, shortcut, sc_stride=1, bn_g=None, bn_h=None, want_lp=False):
    """The squeeze-excite tail of an IR-SE bottleneck in one launch: v NHWC (bf16 / f16) -> (out fp32 = v * gate + shortcut[:, ::s, ::s],
    t_next = out*bn_g + bn_h in v's type or None, out_lp = out in v's type or None); gate = sigmoid(w2 . relu(w1 . mean_hw(v)))."""
    _cuda(v, w1, w2, shortcut, bn_g, bn_h)
    assert v.is_contiguous() and shortcut.is_contiguous() and shortcut.dtype in (v.dtype, torch.float32)
    b, h, w, c = v.shape
    assert shortcut.shape == (b, h * sc_stride, w * sc_stride, c), 'shortcut must be exactly stride x the output size'
    out = torch.empty(b, h, w, c, device=v.device, dtype=torch.float32)
    tn = torch.empty_like(v) if bn_g is not None else None
    lp = torch.empty_like(v) if want_lp else None
    es = _esize(v)
    nbytes = b * h * w * c * (2 * es + shortcut.element_size() + 4 + (0 if tn is None else es) + (0 if lp is None else es))
    with _timed('se_tail', nbytes):
        check(_lib.lib().ood_se_tail(_ptr(v), _ptr(w1), _ptr(w2), _ptr(shortcut), int(sc_stride), _ptr(bn_g), _ptr(bn_h), _ptr(out), _ptr(tn), _ptr(lp),
                                     b, h, w, c, w1.shape[0], _dt(v), int(shortcut.dtype == torch.float32), _stream()), 'se_tail')
    return out, tn, lp


def se_apply(v, tile_sums, w1, w2, shortcut, sc_stride=1, bn_g=None, bn_h=None, want_lp=False):
    """se_tail() with the pooling already done: tile_sums [B, tiles, C, 2] from the conv3x3(..., tile_sums=True) call that wrote v."""
    _cuda(v, tile_sums, w1, w2, shortcut, bn_g, bn_h)
    assert v.is_contiguous() and shortcut.is_contiguous() and shortcut.dtype in (v.dtype, torch.float32)
    b, h, w, c = v.shape
    assert shortcut.shape == (b, h * sc_stride, w * sc_stride, c), 'shortcut must be exactly stride x the output size'
    assert tile_sums.dtype == torch.float32 and tile_sums.is_contiguous() and tile_sums.dim() == 4 and tile_sums.shape[0] == b and \
        tuple(tile_sums.shape[2:]) == (c, 2)
    out = torch.empty(b, h, w, c, device=v.device, dtype=torch.float32)
    tn = torch.empty_like(v) if bn_g is not None else None
    lp = torch.empty_like(v) if want_lp else None
    es = _esize(v)
    nbytes = b * h * w * c * (es + shortcut.element_size() + 4 + (0 if tn is None else es) + (0 if lp is None else es))
    with _timed('se_apply', nbytes):
        check(_lib.lib().ood_se_apply(_ptr(v), _ptr(tile_sums), tile_sums.shape[1], _ptr(w1), _ptr(w2), _ptr(shortcut), int(sc_stride), _ptr(bn_g),
                                      _ptr(bn_h), _ptr(out), _ptr(tn), _ptr(lp), b, h, w, c, w1.shape[0], _dt(v),
                                      int(shortcut.dtype == torch.float32), _stream()), 'se_apply')
    return out, tn, lp


def latent_assemble(heads, stage, avg=None, delta=None):
    """heads fp32 [n_styles, B, D] (the style heads' outputs) -> W+ codes [B, n_styles, D]: w_0 = head_0, w_i = head_0 + head_i up to
    the progressive stage and head_0 beyond (psp_encoders.py:199-214), plus avg [D] and delta [n_styles, D] when given
    (OOD_faceGAN_e4e_arch.py:261)."""
    _cuda(heads, avg, delta)
    assert heads.dtype == torch.float32 and heads.is_contiguous() and heads.dim() == 3
    n, b, dim = heads.shape
    avg, delta = _f32c(avg), _f32c(delta)
    if avg is not None:
        avg = avg.reshape(-1)
        assert avg.numel() == dim
    if delta is not None:
        delta = delta.reshape(-1, dim)
        assert delta.shape[0] == n
    out = torch.empty(b, n, dim, device=heads.device, dtype=torch.float32)
    check(_lib.lib().ood_latent_assemble(_ptr(heads), _ptr(avg), _ptr(delta), _ptr(out), b, n, dim, int(stage), _stream()), 'latent_assemble')
    return out


def alignnet_head_weights(stats, in_w, in_b, w27, w1=None, dtype=torch.bfloat16):
    """-> (wps [B,32,C] `dtype`, bias [B,32] fp32): the AlignNet head's per-sample projection weights with the affine InstanceNorm folded
    in (see ood_b200.h); stats [B,C,2] = {mean, rstd}; w27 [32,C]; w1 [3,C] rides in rows 27..29 when given."""
    _cuda(stats, in_w, in_b, w27, w1)
    b, c, _ = stats.shape
    assert stats.dtype == torch.float32 and stats.is_contiguous() and w27.shape == (32, c) and w27.is_contiguous()
    wps = torch.empty(b, 32, c, device=stats.device, dtype=dtype)
    bias = torch.empty(b, 32, device=stats.device, dtype=torch.float32)
    check(_lib.lib().ood_alignnet_head_weights(_ptr(stats), _ptr(_f32c(in_w)), _ptr(_f32c(in_b)), _ptr(w27), _ptr(_f32c(w1)), _ptr(wps), _ptr(bias),
                                               b, c, _dt(wps), _stream()), 'alignnet_head_weights')
    return wps, bias


def mask_blend(fields, x, gen, want_alpha=True):
    """fields: list of fp32 [B,3,r,r] ascending; x, gen NCHW fp32 [B,3,S,S] -> (out, alpha [B,1,S,S])"""
    _cuda(x, gen, *fields)
    fields = [_f32c(f) for f in fields]
    x, gen = _f32c(x), _f32c(gen)
    b, _, s, _ = x.shape
    out = torch.empty_like(x)
    alpha = torch.empty(b, 1, s, s, device=x.device, dtype=torch.float32) if want_alpha else None
    n = len(fields)
    ptrs = (C.c_void_p * n)(*[f.data_ptr() for f in fields])
    sizes = (C.c_int * n)(*[f.shape[-1] for f in fields])
    work = b * (3 * 3 * s * s * 4 + (s * s * 4 if want_alpha else 0) + sum(f.shape[-1] ** 2 * 4 for f in fields))
    with _timed('mask_blend', work):
        check(_lib.lib().ood_mask_blend(ptrs, sizes, n, _ptr(x), _ptr(gen), _ptr(out), _ptr(alpha), b, s, _stream()), 'mask_blend')
    return out, alpha


def field_step_bwd(z, prev, coarse, gacc, scale, taps=None):
    """Backward of field_step (plain form): -> (gz, gprev or None, gcoarse or None), all fp32."""
    _cuda(z, prev, coarse, gacc)
    z, prev, coarse, gacc = _f32c(z), _f32c(prev), _f32c(coarse), _f32c(gacc)
    b, _, r, _ = z.shape
    rc = coarse.shape[-1] if coarse is not None else 0
    ws, gz = torch.empty_like(z), torch.empty_like(z)
    gprev = torch.empty_like(z) if prev is not None else None
    gcoarse = torch.zeros_like(coarse) if coarse is not None else None
    t = (C.c_float * 4)(*(taps or fir_taps()))
    check(_lib.lib().ood_field_step_bwd(_ptr(z), _ptr(prev), _ptr(coarse), _ptr(gacc), t, float(scale), b, r, rc, _ptr(ws), _ptr(gz),
                                        _ptr(gprev), _ptr(gcoarse), _stream()), 'field_step_bwd')
    return gz, gprev, gcoarse


def warp_mix_bwd(gen, field, gout):
    """Backward of warp_mix: gen, gout NHWC [B,H,W,C]; field fp32 [B,3,H,W] -> (ggen fp32 NHWC, gfield fp32 [B,3,H,W])."""
    _cuda(gen, field, gout)
    assert gen.is_contiguous() and gout.is_contiguous() and gen.shape == gout.shape and gen.dtype == gout.dtype
    b, h, w, c = gen.shape
    ggen = torch.zeros(b, h, w, c, device=gen.device, dtype=torch.float32)
    gfield = torch.zeros(b, 3, h, w, device=gen.device, dtype=torch.float32)
    check(_lib.lib().ood_warp_mix_bwd(_ptr(gen), _ptr(_f32c(field)), _ptr(gout), _ptr(ggen), _ptr(gfield), b, h, w, c, _dt(gen),
                                      _stream()), 'warp_mix_bwd')
    return ggen, gfield


def mask_blend_bwd(fields, x, gen, gout, want_gx=True, want_ggen=True, deterministic=True):
    """Backward of mask_blend: fields list of fp32 [B,3,r,r]; x, gen, gout fp32 [B,3,S,S] -> (gx, ggen, [gfields]).
    deterministic (default): two-stage gather form with a [n,B,S,S] fp32 workspace; False: the atomic form."""
    _cuda(x, gen, gout, *fields)
    fields = [_f32c(f) for f in fields]
    x, gen, gout = _f32c(x), _f32c(gen), _f32c(gout)
    b, _, s, _ = x.shape
    gx = torch.empty_like(x) if want_gx else None
    ggen = torch.empty_like(x) if want_ggen else None
    gfields = [torch.zeros_like(f) for f in fields]
    n = len(fields)
    ptrs = (C.c_void_p * n)(*[f.data_ptr() for f in fields])
    gptrs = (C.c_void_p * n)(*[g.data_ptr() for g in gfields])
    sizes = (C.c_int * n)(*[f.shape[-1] for f in fields])
    ws = torch.empty(n, b, s, s, device=x.device, dtype=torch.float32) if deterministic else None
    check(_lib.lib().ood_mask_blend_bwd(ptrs, gptrs, sizes, n, _ptr(x), _ptr(gen), _ptr(gout), _ptr(gx), _ptr(ggen), _ptr(ws), b, s, _stream()),
          'mask_blend_bwd')
    return gx, ggen, gfields


def conv_scale(cin, k):
    return 1.0 / math.sqrt(cin * k * k)


# ----------------------------------------------------------------------------------------------- differentiable-path blocks
def affine2(x1, x2=None, a=None, b=None, c=None, out=None, channels=None, off1=0, off2=0, off_out=0):
    """out[..., off_out:off_out+C] = a[b,c]*x1[..., off1:off1+C] + b[b,c]*x2[..., off2:off2+C] + c[b,c] on NHWC tensors (a, b, c: fp32
    [B,C] or None = 1, 1, 0).  `out` defaults to a new [B,H,W,C] tensor; slices let one call write half of a concatenation."""
    _cuda(x1, x2, a, b, c, out)
    assert x1.is_contiguous() and (x2 is None or (x2.is_contiguous() and x2.dtype == x1.dtype and x2.shape[:3] == x1.shape[:3]))
    bsz, h, w, p1 = x1.shape
    C_ = channels if channels is not None else p1 - off1
    if out is None:
        out = torch.empty(bsz, h, w, C_ + off_out, device=x1.device, dtype=x1.dtype)
    assert out.is_contiguous() and out.dtype == x1.dtype and out.shape[:3] == x1.shape[:3]
    for t in (a, b, c):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == (bsz, C_))
    check(_lib.lib().ood_nhwc_affine2(_ptr(x1), p1, off1, _ptr(x2), 0 if x2 is None else x2.shape[3], off2, _ptr(a), _ptr(b), _ptr(c), _ptr(out),
                                      out.shape[3], off_out, bsz, h * w, C_, _dt(x1), _stream()), 'nhwc_affine2')
    return out


def prelu(x, slope, g=None):
    """NHWC x, slope fp32 [C]: PReLU(x); with g: the backward g * (x > 0 ? 1 : slope)."""
    _cuda(x, slope, g)
    assert x.is_contiguous() and (g is None or (g.is_contiguous() and g.dtype == x.dtype and g.shape == x.shape))
    out = torch.empty_like(x)
    check(_lib.lib().ood_prelu(_ptr(x), _ptr(g), _ptr(_f32c(slope)), _ptr(out), x.numel() // x.shape[-1], x.shape[-1], _dt(x), _stream()), 'prelu')
    return out


def tap_gather(g, cp=32, dtype=torch.float32):
    """Adjoint of tap_sum: g fp32 NCHW [B,3,H,W] -> NHWC [B,H,W,cp] `dtype`, channel 3*tap + colour (see ood_b200.h)."""
    _cuda(g)
    g = _f32c(g)
    b, _, h, w = g.shape
    out = torch.empty(b, h, w, cp, device=g.device, dtype=dtype)
    check(_lib.lib().ood_tap_gather(_ptr(g), _ptr(out), b, h, w, cp, _dt(out), _stream()), 'tap_gather')
    return out


def conv_wgrad(g, x, taps=9, cout=None, cin=None):
    """Weight gradient of a shared-weight convolution: g NHWC bf16 [B,H,W,Co_p] (gradient of the output), x NHWC bf16 [B,H,W,Ci_p] (its
    input) -> fp32 [cout, cin, k, k] (k = 3 for taps 9, pad 1; 1 for taps 1), PyTorch's Conv2d weight layout.  tcgen05, deterministic."""
    _cuda(g, x)
    assert g.is_contiguous() and x.is_contiguous() and g.dtype == torch.bfloat16 and x.dtype == torch.bfloat16 and g.shape[:3] == x.shape[:3]
    b, h, w, co_p = g.shape
    ci_p = x.shape[3]
    cout, cin = cout or co_p, cin or ci_p
    nws = _lib.lib().ood_conv_wgrad_workspace(b, h, w, ci_p, co_p, taps) // 4
    if nws <= 0:
        raise RuntimeError(f'ood_gan_inversion_b200: conv_wgrad envelope: cout % 128 == 0, cin % 64 == 0, 64-pixel patches (got {co_p}, {ci_p}, {h}x{w})')
    ws = torch.empty(nws, device=g.device, dtype=torch.float32)
    k = 3 if taps == 9 else 1
    out = torch.empty(cout, cin, k, k, device=g.device, dtype=torch.float32)
    with _timed('conv_wgrad', 2.0 * b * h * w * co_p * ci_p * taps):
        check(_lib.lib().ood_conv_wgrad(_ptr(g), _ptr(x), _ptr(ws), _ptr(out), b, h, w, ci_p, co_p, taps, cin, cout, _stream()), 'conv_wgrad')
    return out


# ---- device guard over every public wrapper (a model on cuda:1 in a process whose current device is cuda:0) ----
def _install_device_guard():
    import types
    g = globals()
    skip = {'profile_begin', 'profile_end', 'fir_taps'}
    for name, obj in list(g.items()):
        if isinstance(obj, types.FunctionType) and obj.__module__ == __name__ and not name.startswith('_') and name not in skip:
            g[name] = _device_guard(obj)


_install_device_guard()
