"""Drop-in for the reference's Spatial Alignment & Masking Module (src/ops/SAMM/helpers.py).

Same module tree and state-dict keys (`alignment.body.body.{0,1}.{res_layer,shortcut_layer}.*`, `alignment.blur.kernel`,
`weight`, `noiseInj.weight`).  The field head (tanh/sigmoid + FIR blur + accumulate/PRM/clip + coarse-level bicubic PRM),
the flow warp + alpha mix and (in arch.py) the mask compose + blend are single fused sm_100a kernels; the AlignNet
convolution stacks run on the package's tcgen05 convolution in shared-weight mode (AlignNet.raw_nhwc; SURVEY.md section
8(f) rank 1), with the `enc`-only half of the first convolution computed once per level and reused by every cycle.
"""
import os

import torch
from torch import nn
from torch.nn import functional as F

from . import kernels as K
from . import stylegan as sg
from .stylegan import Blur, NoiseInjection


# Levels whose first AlignNet convolution is split across the alignment cycles (AlignNet.raw_nhwc): below this channel count
# the half-K convolutions are bound by their epilogue and operand feed rather than by the tensor pipe, and the seed's extra
# traffic costs more than the saved MACs (scripts/seed_bench.py on B200: C=128 at 256 px, whole 1.83 ms vs split 2.0 ms per
# two cycles; C=256 at 128 px 1.66 vs 1.42; C=512 at 64 px 1.58 vs 1.24).
_SPLIT_MIN_C = int(os.environ.get('OOD_SPLIT_MIN_C', 256))
# A/B switches for the measurements in profiles/README.md (defaults = the fast route)
_CONV_STATS = os.environ.get('OOD_CONV_STATS', '1') != '0'      # InstanceNorm statistics from the convolution's epilogue
_FOLD_SHORTCUT = os.environ.get('OOD_FOLD_SHORTCUT', '1') != '0'  # 1x1 shortcut in the projection's spare output channels


def BN(depth, bn=True):
    """src/ops/e4e/encoders/helpers.py:93-99"""
    if bn == 'InstanceNorm':
        return nn.InstanceNorm2d(depth, affine=True)
    if bn == 'BatchNorm' or bn is True:
        return nn.BatchNorm2d(depth)
    return nn.Identity()


class bottleneck_IR(nn.Module):
    """src/ops/e4e/encoders/helpers.py:426-448"""

    def __init__(self, in_channel, depth, stride, bn=True, bias=False):
        super().__init__()
        if in_channel == depth:
            self.shortcut_layer = nn.MaxPool2d(1, stride)
        else:
            self.shortcut_layer = nn.Sequential(nn.Conv2d(in_channel, depth, (1, 1), stride, bias=bias), BN(depth, bn=bn))
        self.res_layer = nn.Sequential(BN(in_channel, bn=bn), nn.Conv2d(in_channel, depth, (3, 3), (1, 1), 1, bias=bias),
                                       nn.PReLU(depth), nn.Conv2d(depth, depth, (3, 3), stride, 1, bias=bias), BN(depth, bn=bn))

    def forward(self, x):
        if isinstance(self.shortcut_layer, nn.MaxPool2d):       # MaxPool2d(1, s) == strided subsampling: no kernel needed
            s = self.shortcut_layer.stride
            return self.res_layer(x) + (x if s == 1 else x[:, :, ::s, ::s])
        return self.res_layer(x) + self.shortcut_layer(x)


def scaleNshiftBlock(in_chn, out_chn, norm_type=False, bias=False):
    """src/ops/SAMM/helpers.py:58-60"""
    return nn.Sequential(bottleneck_IR(in_chn, in_chn, 1, bn=norm_type, bias=bias),
                         bottleneck_IR(in_chn, out_chn, 1, bn=norm_type, bias=bias))


def new_PRM(x, y, **kwargs):
    """src/ops/SAMM/helpers.py:62-77 (torch form, API parity; the hot path uses ood_field_step)."""
    if x.shape[-2:] != y.shape[-2:]:
        x = F.interpolate(x, size=y.shape[-2:], mode='bicubic', align_corners=True)
    return y * x + x * (1 - x)


class AlignNet(nn.Module):
    """src/ops/SAMM/helpers.py:85-109"""

    def __init__(self, in_chn, out_chn=3, scale=1., blur_kernel=[1, 3, 3, 1], **kwargs):
        super().__init__()
        self.norm = nn.InstanceNorm2d(in_chn)
        self.body = scaleNshiftBlock(in_chn * 2, out_chn, 'InstanceNorm', kwargs.get('bias', False))
        self.tanh = nn.Tanh()
        self.sigmoid = nn.Sigmoid()
        self.scale = scale
        self.diff_fAndg = kwargs.get('diff_fAndg', True)

    def raw(self, source, target):
        """Pre-activation 3-channel output (fp32); the heads are fused into ood_field_step."""
        bf16 = source.dtype == torch.bfloat16
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False), \
                torch.autocast('cuda', dtype=torch.bfloat16, enabled=bf16):
            source, target = self.norm(source), self.norm(target)
            z = torch.cat([source - target, target] if self.diff_fAndg else [source, target], dim=1)
            z = self.body(z)
        return z.float()

    # ---- fused NHWC route: every 2C-channel pass runs on this package's kernels (shared-weight tcgen05 / SIMT conv with
    # a PReLU epilogue, one-pass pair statistics, re-derived residual); only the 3-channel fp32 tail stays torch ops.
    def fused_ok(self):
        convs = [self.body[0].res_layer[1], self.body[0].res_layer[3], self.body[1].res_layer[1], self.body[1].res_layer[3],
                 self.body[1].shortcut_layer[0]]
        c2 = convs[0].weight.shape[0]
        return self.diff_fAndg and all(c.bias is None for c in convs) and c2 % sg._granule() == 0 and \
            isinstance(self.body[0].res_layer[0], nn.InstanceNorm2d) and isinstance(self.body[1].res_layer[4], nn.InstanceNorm2d) and \
            isinstance(self.body[1].shortcut_layer[1], nn.InstanceNorm2d)

    def _packed(self):
        b0, b1 = self.body[0], self.body[1]
        params = [b0.res_layer[1].weight, b0.res_layer[3].weight, b1.res_layer[1].weight, b1.shortcut_layer[0].weight]
        key = (sg.get_precision(), params[0].device) + tuple((p._version, p.data_ptr()) for p in params)
        hit = getattr(self, '_pk', None)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                dt, cim = sg._act_dtype(), sg.get_precision() == 'fp32'
                g = sg._granule()
                # 2C -> 3 head as a per-pixel projection onto the 27 (tap, colour) weights (ood_tap_sum): row 3*t + k
                wc = b1.res_layer[1].weight.detach().float()                                        # [3, 2C, 3, 3]
                w27 = torch.zeros(32, wc.shape[1], device=wc.device)
                w27[:27] = wc.permute(2, 3, 0, 1).reshape(27, -1)
                wa = b0.res_layer[1].weight.detach()
                half = wa.shape[1] // 2
                # first convolution split along its input channels: [IN(cur)-IN(enc)] half | [IN(enc)] half (see raw_nhwc)
                split = dict(wa_lo=K.pack_conv_weight(wa[:, :half].contiguous(), dt, cim),
                             wa_hi=K.pack_conv_weight(wa[:, half:].contiguous(), dt, cim)) if not cim else {}
                pk = dict(wa=K.pack_conv_weight(wa, dt, cim), **split,
                          wb=K.pack_conv_weight(b0.res_layer[3].weight.detach(), dt, cim),
                          wc=K.pack_conv1x1_weight(w27, dt, cim), cp=32, w27=w27.contiguous(),
                          w1=b1.shortcut_layer[0].weight.detach().float().reshape(3, -1).contiguous())
            self._pk = (key, pk)
            hit = self._pk
        return hit[1]

    def raw_nhwc(self, cur, enc, fold=False, carry=None):
        """cur, enc: NHWC [B,R,R,C] in the pipeline's storage type -> pre-activation field [B,3,R,R] fp32.
        fold=True returns (r2, shortcut, coef) instead: the field is r2*coef[...,0] + shortcut*coef[...,1] + coef[...,2], which
        ood_field_step applies on load (the two closing InstanceNorms and the residual sum never take a pass of their own).
        carry: a dict the caller keeps across the alignment cycles of ONE `enc` (SPM_Warp.forward_nhwc).  The first
        convolution is linear in its input cat[INaff(IN(cur)-IN(enc)), INaff(IN(enc))]; its second half depends on `enc`
        alone, so that half of the contraction (fp32 accumulators) is computed in the first cycle and seeds the accumulators
        of the [IN(cur)-IN(enc)] half in every cycle: 3.5 instead of 4 wide convolutions per two-cycle level."""
        b0, b1 = self.body[0], self.body[1]
        pk = self._packed()
        f = lambda p: p.detach().float().contiguous()
        b, r, _, c = cur.shape
        impl = sg._impl()
        eps = self.norm.eps
        st6 = K.in_stats(cur, enc, eps)
        split = impl == 0 and carry is not None and c % 64 == 0 and c >= _SPLIT_MIN_C
        if split:
            seed = carry.get('seed')
            lo, hi = K.alignnet_front_split(cur, enc, st6, f(b0.res_layer[0].weight), f(b0.res_layer[0].bias), want_hi=seed is None)
            if seed is None:
                seed, _ = K.conv3x3(hi, pk['wa_hi'], 2 * c, impl=0, out_f32=True, tiled=True)
                carry['seed'] = seed
            x, _ = K.conv3x3(lo, pk['wa_lo'], 2 * c, impl=0, prelu=f(b0.res_layer[2].weight), acc_in=seed, tiled=True)
        else:
            x = K.alignnet_front(cur, enc, st6, f(b0.res_layer[0].weight), f(b0.res_layer[0].bias))
            x, _ = K.conv3x3(x, pk['wa'], 2 * c, impl=impl, prelu=f(b0.res_layer[2].weight))
        if impl == 0 and _CONV_STATS and K.conv3x3_stats_ok(x, 2 * c):      # InstanceNorm statistics of the output from the convolution's epilogue
            x, _, st_x = K.conv3x3(x, pk['wb'], 2 * c, impl=0, stats_eps=b0.res_layer[4].eps)
        else:
            x, _ = K.conv3x3(x, pk['wb'], 2 * c, impl=impl)
            st_x = K.in_stats(x, None, b0.res_layer[4].eps)
        nvec = c // (8 if cur.dtype == torch.bfloat16 else 4)
        if nvec <= 256 and 256 % nvec == 0:
            out0, st = K.alignnet_res0_stats(x, st_x, f(b0.res_layer[4].weight), f(b0.res_layer[4].bias), cur, enc, st6, eps)
        else:
            out0 = K.alignnet_res0(x, st_x, f(b0.res_layer[4].weight), f(b0.res_layer[4].bias), cur, enc, st6)
            st = K.in_stats(out0, None, eps)
        if impl == 0:
            # InstanceNorm(out0) folded into the projection: W27 . (g*out0 + h) = (W27 diag(g_b)) . out0 + W27 . h_b -- one weight
            # set per sample (grouped form, groups = batch) and a per-sample bias, so the normalised copy is never written
            # rows 27..29 of the projection: the bottleneck's 1x1 shortcut convolution on the UN-normalised out0 (no g, no bias),
            # so out0 is read once for both branches; ood_tap_sum_shortcut hands those channels back as planes
            wps, hb = K.alignnet_head_weights(st, b1.res_layer[0].weight, b1.res_layer[0].bias, pk['w27'], pk['w1'] if _FOLD_SHORTCUT else None)
            x, _ = K.conv3x3(out0, wps, pk['cp'], transposed=4, impl=0, out_f32=True, groups=b, bias=hb)       # [groups][1 tap][Co][Ci]
            if _FOLD_SHORTCUT:
                res, sc = K.tap_sum(x, shortcut=True)
            else:
                res = K.tap_sum(x)
                sc = K.torgb(out0, pk['w1'].unsqueeze(0).expand(b, -1, -1).contiguous(), torch.zeros(3, device=cur.device))
        else:
            x = K.in_apply(out0, st, f(b1.res_layer[0].weight), f(b1.res_layer[0].bias))
            x, _ = K.conv3x3(x, pk['wc'], pk['cp'], transposed=4, impl=impl, out_f32=True)
            res = K.tap_sum(x)
            zero = torch.zeros(3, device=cur.device)
            sc = K.torgb(out0, pk['w1'].unsqueeze(0).expand(b, -1, -1).contiguous(), zero)      # 1x1 conv 2C -> 3, fp32 NCHW
        n_res, n_sc = b1.res_layer[4], b1.shortcut_layer[1]
        r2, coef = K.alignnet_tail(res, sc, f(b1.res_layer[2].weight), f(b1.res_layer[3].weight), f(n_res.weight), f(n_res.bias),
                                   f(n_sc.weight), f(n_sc.bias), n_res.eps)
        if fold:
            return r2, sc, coef
        return r2 * coef[:, :, 0, None, None] + sc * coef[:, :, 1, None, None] + coef[:, :, 2, None, None]

    def raw_diff(self, cur, enc):
        """Differentiable form of raw_nhwc (NHWC cur, enc -> pre-activation field [B,3,R,R] fp32) for the gradient path: the same
        arithmetic, un-fused, as a graph of diff_ops Functions (every 2C-channel pass forward and backward is a kernel of this
        library; the 3-channel fp32 tail -- PReLU(3), conv3x3(3 -> 3), two InstanceNorm(3) -- is plain torch ops on [B,3,R,R]).
        Gradients flow to `cur` (and to `enc` if it requires grad); see diff_ops for the weight-gradient status."""
        from . import diff_ops as D
        if not self.fused_ok():
            raise NotImplementedError('ood_gan_inversion_b200: the differentiable AlignNet needs the default configuration (diff_fAndg, no biases)')
        b0, b1 = self.body[0], self.body[1]
        eps = self.norm.eps
        c_hat = D.inst_norm(cur, None, None, eps)
        with torch.set_grad_enabled(enc.requires_grad):
            e_hat = D.inst_norm(enc, None, None, eps)
        z = D.cat(D.sub(c_hat, e_hat), e_hat)                                    # SAMM/helpers.py:96-101
        n0, n4 = b0.res_layer[0], b0.res_layer[4]
        u = D.inst_norm(z, n0.weight, n0.bias, n0.eps)
        h = D.prelu(D.conv(u, b0.res_layer[1].weight), b0.res_layer[2].weight)
        v = D.inst_norm(D.conv(h, b0.res_layer[3].weight), n4.weight, n4.bias, n4.eps)
        out0 = D.add(z, v)                                                      # bottleneck_IR with the identity shortcut
        m0 = b1.res_layer[0]
        res = D.head27(D.inst_norm(out0, m0.weight, m0.bias, m0.eps), b1.res_layer[1].weight)      # fp32 [B,3,R,R]
        sc = D.shortcut3(out0, b1.shortcut_layer[0].weight)
        n_res, n_sc = b1.res_layer[4], b1.shortcut_layer[1]
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            r = F.conv2d(F.prelu(res, b1.res_layer[2].weight.float()), b1.res_layer[3].weight.float(), padding=1)
        r = F.instance_norm(r, weight=n_res.weight.float(), bias=n_res.bias.float(), eps=n_res.eps)
        s = F.instance_norm(sc, weight=n_sc.weight.float(), bias=n_sc.bias.float(), eps=n_sc.eps)
        return s + r

    def forward(self, source, target, **kwargs):
        z = self.raw(source, target)
        return torch.cat([self.tanh(z[:, 0:1]) * self.scale, self.tanh(z[:, 1:2]) * self.scale, self.sigmoid(z[:, 2:])], dim=1)


def weight_init(m):
    """src/ops/SAMM/helpers.py:79-83"""
    if isinstance(m, nn.Conv2d):
        nn.init.constant_(m.weight, 0)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)


class SPM_Warp(nn.Module):
    """src/ops/SAMM/helpers.py:111-179"""

    def __init__(self, in_chn, scale=0.1, style_dim=512, blur_kernel=[1, 3, 3, 1], cycle_align=1, **kwargs):
        super().__init__()
        self.body = AlignNet(in_chn, 3, scale=scale, style_dim=style_dim, blur_kernel=blur_kernel, **kwargs)
        self.body.apply(weight_init)
        self.scale = scale
        self.cycle_align = cycle_align
        self.weight_init()
        self.blur = Blur(kernel=blur_kernel, pad=(2, 1))

    def weight_init(self):
        for module in self.modules():
            if isinstance(module, (nn.Conv2d, sg.ModulatedConv2d, nn.Linear)):
                nn.init.xavier_normal_(module.weight)

    def forward_nhwc(self, source, target_nhwc, aligned=None):
        """source: encoder features, logical [B,C,R,R] (any memory format); target_nhwc: generator features
        [B,R,R,C] in the pipeline's storage type; aligned: coarser level's field or None.
        Returns (aligned features NHWC, field fp32 [B,3,R,R])."""
        if self.blur.taps is None:
            raise NotImplementedError('ood_gan_inversion_b200: SPM_Warp needs a 4-tap blur kernel')
        fused = self.body.fused_ok()
        src = source.to(target_nhwc.dtype)
        if fused:
            src = src.permute(0, 2, 3, 1).contiguous()               # encoder features -> NHWC once per level
        cur, acc = target_nhwc, None
        carry = {}                      # per-call state shared by the cycles (the enc-only part of the first convolution)
        for k in range(self.cycle_align):
            if fused:
                z, z2, coef = self.body.raw_nhwc(cur, src, fold=True, carry=carry if self.cycle_align > 1 else None)
            else:
                z, z2, coef = self.body.raw(cur.permute(0, 3, 1, 2), src), None, None      # NHWC storage viewed as channels_last NCHW
            last = k == self.cycle_align - 1
            acc = K.field_step(z, acc, aligned if last else None, self.scale, self.blur.taps, z2=z2, coef=coef)
            cur = K.warp_mix(target_nhwc, acc)
        return cur, acc

    def forward_nhwc_diff(self, source, target_nhwc, aligned=None):
        """forward_nhwc for the gradient path (SAMM/helpers.py:149-179 under autograd): differentiable in the generator features
        `target_nhwc`, in the coarser level's field `aligned` (its alpha channel) and in `source` if it requires grad.  Forward and
        backward of the gather / field steps are ood_warp_mix(_bwd) / ood_field_step(_bwd) (samm_grad)."""
        from . import samm_grad
        if self.blur.taps is None:
            raise NotImplementedError('ood_gan_inversion_b200: SPM_Warp needs a 4-tap blur kernel')
        src = source.to(target_nhwc.dtype).permute(0, 2, 3, 1).contiguous()
        cur, acc = target_nhwc, None
        for k in range(self.cycle_align):
            z = self.body.raw_diff(cur, src)
            last = k == self.cycle_align - 1
            acc = samm_grad.field_step(z, acc, aligned if last else None, self.scale)
            cur = samm_grad.warp_mix(target_nhwc, acc)
        return cur, acc

    def forward(self, source, target, style=None, aligned=None):
        """Reference contract: NCHW fp32 in, (aligned_target NCHW fp32, field) out."""
        t = K.nchw_to_nhwc(target, None, sg._act_dtype())
        cur, acc = self.forward_nhwc(source, t, aligned)
        return K.nhwc_to_nchw(cur), acc


class StyledscaleNshfitBlock(nn.Module):
    """src/ops/SAMM/helpers.py:182-216 (default configuration: identity btn1, alignment on)."""

    def __init__(self, in_chn, out_chn, style_dim, alignment=True, btn='style_bottleneck_IR', **kwargs):
        super().__init__()
        if btn is not None:
            raise NotImplementedError("ood_gan_inversion_b200: only mod_btn=None (identity btn1) is implemented; no shipped "
                                      'config sets mod_btn (SURVEY appendix B.5)')
        self.btn1 = lambda x, y: x
        out_chn = in_chn
        if alignment:
            self.alignment = SPM_Warp(out_chn, **kwargs)
        else:
            self.alignment = None
        self.weight = nn.Parameter(torch.ones(1), requires_grad=False)
        self.noiseInj = NoiseInjection()

    def forward_nhwc(self, x, image_nhwc, aligned=None):
        if self.alignment is None:
            return x, None
        if torch.is_grad_enabled() and (image_nhwc.requires_grad or (aligned is not None and aligned.requires_grad) or x.requires_grad):
            return self.alignment.forward_nhwc_diff(x, image_nhwc, aligned)
        return self.alignment.forward_nhwc(x, image_nhwc, aligned)

    def forward(self, x, styles, **kwargs):
        gen_feat = kwargs.get('image', None)
        assert gen_feat is not None
        if kwargs.get('transform', None) is not None:
            x = kwargs['transform'](x)
        if self.alignment is None:
            return x, None
        return self.alignment(x, gen_feat, styles, kwargs.get('aligned', None))
