"""bf16 inference path of the E4E encoder on this library's kernels (NHWC activations).

Reference module: src/ops/e4e/encoders/psp_encoders.py:34-56,125-216 and helpers.py:59-76,476-501 (eval mode,
e4e_arch.py:256-258).  The trunk's 48 stride-1 / stride-2 3x3 convolutions and the 98 stride-2 convolutions of the 18
GradualStyleBlocks run on the tcgen05 implicit-GEMM kernel (`ood_conv3x3`, forms 0 and 3) with their PReLU / folded
BatchNorm bias / LeakyReLU fused into the epilogue; the squeeze-excite gate and the residual sum (which also emits the
next block's BatchNorm) are `ood_se_gate` / `ood_se_residual`; the 3->64 input convolution runs with its input channels
zero-padded to 32, the lateral 1x1 convolutions and the heads' closing EqualLinear on the (grouped) 1x1 form, the three 1x1
stride-2 shortcut convolutions on its strided form; the residual stream is kept in fp32.  No cuDNN / cuBLAS call is left.

Built from (and numerically checked against) `encoder.Encoder4Editing`; eval-mode BatchNorms are folded in fp32.
"""
import os

import torch
from torch import nn
from torch.nn import functional as F

from . import kernels as K

_SE_SUMS = os.environ.get('OOD_SE_SUMS', '1') != '0'         # A/B switch: pooled sums from the second convolution's epilogue + ood_se_apply (stages 2-4)
_SE_SUMS_MIN_C = int(os.environ.get('OOD_SE_SUMS_MIN_C', 256))   # 128 channels (64 px): the sums force 128-wide tiles on the convolution, +11 us for -10 us
_SE_FUSED = os.environ.get('OOD_SE_FUSED', '1') != '0'       # A/B switch: the one-launch squeeze-excite tail (ood_se_tail) vs the four-launch route
# Storage type of the encoder's activations and weights on the tensor-core path.  IEEE half, not bf16: every activation here is
# behind an eval-mode BatchNorm (O(1) values, far inside the half range; conversions saturate), the tcgen05 pipe runs f16 and
# bf16 at the same rate, and 11 significant bits instead of 8 cut the W+ error against the fp32 reference from 0.9 % to ~0.1 %
# rel-L2 -- with bf16 the encoder's latents were the largest single term of the pipeline's image error (scripts/diag_bf16_budget.py:
# 0.0148 max-abs all-bf16 against 0.0091 with fp32 latents, batch 8).  OOD_ENCODER_DTYPE=bf16 restores the old route for A/B runs.
ENC_DT = torch.bfloat16 if os.environ.get('OOD_ENCODER_DTYPE', 'f16') == 'bf16' else torch.float16


def _bn_affine(bn):
    g = bn.weight.detach().float() * torch.rsqrt(bn.running_var.detach().float() + bn.eps)
    return g.contiguous(), (bn.bias.detach().float() - bn.running_mean.detach().float() * g).contiguous()


def _pack(w):
    return K.pack_conv_weight(w.float().contiguous(), ENC_DT, False)


def _nhwc(t):
    """NCHW channels_last tensor -> NHWC contiguous view (no copy when already channels_last)."""
    return t.permute(0, 2, 3, 1).contiguous()


def _nchw(t):
    """NHWC contiguous -> NCHW view in channels_last memory format (no copy)."""
    return t.permute(0, 3, 1, 2)


class _Block:
    def __init__(self, blk):
        bn1, conv1, prelu, conv2, bn2, se = list(blk.res_layer.children())
        self.stride = conv2.stride[0]
        self.depth = conv1.out_channels
        self.bn1 = _bn_affine(bn1)
        self.w1 = _pack(conv1.weight.detach())
        self.slope = prelu.weight.detach().float().contiguous()
        g2, h2 = _bn_affine(bn2)
        self.w2 = _pack(conv2.weight.detach().float() * g2.reshape(-1, 1, 1, 1))
        self.b2 = h2
        self.se1 = se.fc1.weight.detach().float().reshape(se.fc1.out_channels, -1).contiguous()
        self.se2 = se.fc2.weight.detach().float().reshape(se.fc2.out_channels, -1).contiguous()
        self.shortcut = None
        if isinstance(blk.shortcut_layer, nn.Sequential):
            conv, bn = list(blk.shortcut_layer.children())
            g, h = _bn_affine(bn)
            # Conv2d(in, depth, 1, stride, bias=False) + BatchNorm2d (helpers.py:483-486), folded: the 1x1 (stride-2) form of the
            # tcgen05 kernel reads every second pixel of every second row through the TMA element strides
            if conv.stride[0] not in (1, 2):
                raise NotImplementedError('FastEncoder: shortcut stride must be 1 or 2')
            self.shortcut = (K.pack_conv1x1_weight(conv.weight.detach().float() * g.reshape(-1, 1, 1, 1), ENC_DT, False), h,
                             conv.out_channels, 6 if conv.stride[0] == 2 else 4)


class _HeadGroup:
    """All GradualStyleBlocks that read the same feature map (3 coarse / 4 middle / 11 fine heads): their convolution
    chains have identical shapes, so every depth is ONE grouped launch (`groups` = heads) instead of one tiny launch per
    head -- at 16 px and below a single head's convolution is a few CTAs streaming 4.7 MB of weights."""

    def __init__(self, heads):
        self.n = len(heads)
        chains = [[m for m in h.convs.children() if isinstance(m, nn.Conv2d)] for h in heads]
        self.depth = len(chains[0])
        self.out_c = heads[0].out_c
        self.layers = []
        for d in range(self.depth):
            w = torch.cat([_pack(c[d].weight.detach()) for c in chains], 0)                       # [n*9, Co, Ci]
            b = torch.stack([c[d].bias.detach().float() for c in chains]).contiguous()           # [n, Co]
            slope = torch.full_like(b, 0.01)                                                     # nn.LeakyReLU()
            self.layers.append((w, b, slope, chains[0][d].out_channels))
        # the closing EqualLinear of every head (psp_encoders.py:50-55) as ONE grouped 1x1 launch: pack [n][out][in]
        self.lw = torch.cat([K.pack_conv1x1_weight(h.linear.weight.detach().float() * h.linear.scale, ENC_DT, False)
                             for h in heads], 0).contiguous()
        self.lb = torch.stack([h.linear.bias.detach().float() * h.linear.lr_mul for h in heads]).contiguous()        # [n, out]

    def __call__(self, x, out=None):
        """x NHWC [B,S,S,512] -> [n, B, 512] fp32 latents (written into `out`, a contiguous [n, B, 512] fp32 slice, when given)"""
        b = x.shape[0]
        for d, (w, bias, slope, co) in enumerate(self.layers):
            x, _ = K.conv3x3(x, w, co, transposed=3, bias=bias, prelu=slope, tag='encoder_conv', groups=self.n, in_shared=(d == 0))
        y, _ = K.conv3x3(x.reshape(self.n * b, 1, 1, self.out_c), self.lw, self.out_c, transposed=4, bias=self.lb, tag='encoder_conv',
                         groups=self.n, out_f32=True, out=None if out is None else out.view(self.n * b, 1, 1, self.out_c))
        return y.reshape(self.n, b, self.out_c)


class FastEncoder:
    """Callable with the signature of Encoder4Editing.forward(x, return_feats=...) for bf16 channels_last inputs."""

    @torch.no_grad()
    def __init__(self, enc):
        conv, bn, prelu = list(enc.input_layer.children())
        g, h = _bn_affine(bn)
        # input layer (3 -> 64, BatchNorm folded, PReLU in the epilogue) on the tcgen05 kernel: input channels zero-padded to 32
        w0 = torch.zeros(conv.out_channels, 32, 3, 3, device=conv.weight.device)
        w0[:, :3] = conv.weight.detach().float() * g.reshape(-1, 1, 1, 1)
        self.first_w, self.first_b, self.first_c = _pack(w0), h, conv.out_channels
        self.first_slope = prelu.weight.detach().float().contiguous()
        self.blocks = [_Block(b) for b in enc.body]
        self.style_count, self.coarse_ind, self.middle_ind = enc.style_count, enc.coarse_ind, enc.middle_ind
        styles = list(enc.styles)
        self.head_groups = [_HeadGroup(styles[:self.coarse_ind]), _HeadGroup(styles[self.coarse_ind:self.middle_ind]),
                            _HeadGroup(styles[self.middle_ind:])]
        self.lat = []                                        # lateral 1x1 convolutions (bias in the epilogue)
        for lat in (enc.latlayer1, enc.latlayer2):
            self.lat.append((K.pack_conv1x1_weight(lat.weight.detach(), ENC_DT, False), lat.bias.detach().float().contiguous(),
                             lat.out_channels))
        self.progressive_stage = enc.progressive_stage

    def _lateral(self, i, x):
        w, b, co = self.lat[i]
        return K.conv3x3(x, w, co, transposed=4, bias=b, tag='encoder_conv')[0]

    @torch.no_grad()
    def __call__(self, x, return_feats=False, thumb=None, avg=None, delta=None, **kwargs):
        """x: fp32 NCHW [B,3,256,256] (the module's input), or thumb: the same image as the first convolution's operand, bf16 NHWC
        [B,256,256,32] with channels 3..31 zero (kernels.thumbnail_nhwc: resize + layout + cast + padding in one pass).
        avg [1,512] / delta [1,18,512]: the arch's latent offsets (e4e_arch.py:261), added inside the W+ assembly kernel."""
        if thumb is not None:
            xp = thumb
            if xp.dtype != ENC_DT or xp.dim() != 4 or xp.shape[-1] != 32 or not xp.is_contiguous():
                raise ValueError(f'FastEncoder: thumb must be a contiguous {ENC_DT} NHWC tensor with 32 channels')
        else:
            if not x.is_cuda:
                raise RuntimeError('ood_gan_inversion_b200 is CUDA-only')
            xp = torch.zeros(x.shape[0], x.shape[2], x.shape[3], 32, device=x.device, dtype=ENC_DT)
            xp[..., :3] = x.permute(0, 2, 3, 1)
        x0, _ = K.conv3x3(xp, self.first_w, self.first_c, bias=self.first_b, prelu=self.first_slope, tag='encoder_conv')
        feats = [_nchw(x0)]
        # The residual stream `cur` is fp32 from the first block on (out_f32): rounding the running sum to bf16 after each of
        # the 24 blocks was the largest single error source of the bf16 pipeline (alpha error 1.2e-2 -> 0.5e-2 without it).
        # t = BN1(cur) of the first block; afterwards every residual pass emits the next block's BN1 itself.
        cur = x0
        _, t = K.se_residual(cur, bn_g=self.blocks[0].bn1[0], bn_h=self.blocks[0].bn1[1], want_out=False)
        taps = {}
        for i, blk in enumerate(self.blocks):
            u, _ = K.conv3x3(t, blk.w1, blk.depth, prelu=blk.slope, tag='encoder_conv')
            form = 3 if blk.stride == 2 else 0
            sums = None
            if _SE_SUMS and blk.depth >= _SE_SUMS_MIN_C and K.conv3x3_stats_ok(u, blk.depth, form):
                # SEModule's global average pool (helpers.py:59-76) rides in this convolution's epilogue: per-tile channel sums
                v, _, sums = K.conv3x3(u, blk.w2, blk.depth, transposed=form, bias=blk.b2, tag='encoder_conv', tile_sums=True)
            else:
                v, _ = K.conv3x3(u, blk.w2, blk.depth, transposed=form, bias=blk.b2, tag='encoder_conv')
            if blk.shortcut is not None:
                # the blocks with a shortcut convolution (3, 7, 21) follow the tapped blocks: their bf16 input is already there
                w_sc, b_sc, c_sc, sform = blk.shortcut
                src = taps[i - 1] if (i - 1) in taps else cur.to(ENC_DT)
                sc, ss = K.conv3x3(src, w_sc, c_sc, transposed=sform, bias=b_sc, tag='encoder_conv')[0], 1
            else:
                sc, ss = cur, blk.stride                                   # MaxPool2d(1, s): strided read
            nxt = self.blocks[i + 1].bn1 if i + 1 < len(self.blocks) else (None, None)
            tap = i in (2, 6, 20, 23)            # tapped blocks: the fp32 stream once more in the storage type, from the same pass
            if sums is not None:
                cur, t, lp = K.se_apply(v, sums, blk.se1, blk.se2, sc, ss, nxt[0], nxt[1], want_lp=tap)
            elif _SE_FUSED:
                # SEModule's global average pool + gate MLP + residual + next BatchNorm (helpers.py:59-76, 494-501): one cluster launch
                cur, t, lp = K.se_tail(v, blk.se1, blk.se2, sc, ss, nxt[0], nxt[1], want_lp=tap)
            else:
                gate = K.se_gate(K.in_stats(v), blk.se1, blk.se2)
                r = K.se_residual(v, gate, sc, ss, nxt[0], nxt[1], out_f32=True, want_lp=tap)
                cur, t, lp = r if tap else (r[0], r[1], None)
            if tap:
                taps[i] = lp
                feats.append(_nchw(lp))
        c1, c2, c3 = taps[6], taps[20], taps[23]
        # psp_encoders.py:199-214: w_i = w_0 + head_i(features); heads beyond the progressive stage repeat w_0
        stage = self.progressive_stage.value
        b = c3.shape[0]
        heads = torch.empty(self.style_count, b, self.head_groups[0].out_c, device=c3.device, dtype=torch.float32)
        self.head_groups[0](c3, out=heads[:self.coarse_ind])                   # [3, B, 512]
        if stage >= self.coarse_ind:
            p2 = K.bicubic_up_add(c3, self._lateral(0, c2))
            self.head_groups[1](p2, out=heads[self.coarse_ind:self.middle_ind])
        if stage >= self.middle_ind:
            p1 = K.bicubic_up_add(p2, self._lateral(1, c1))
            self.head_groups[2](p1, out=heads[self.middle_ind:])
        w = K.latent_assemble(heads, min(stage, self.style_count - 1), avg, delta)
        return (w, feats) if return_feats else w
