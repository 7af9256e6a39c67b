"""Inference-time copy of the E4E encoder for the bf16 path: eval-mode BatchNorms that FOLLOW a convolution are folded into
it (exactly, in fp32), and every remaining parameter is stored once in bf16 / channels_last, so a forward neither
re-casts ~270 M fp32 weights (autocast does that on every call) nor runs the folded normalisation passes.

The encoder is in front of the hot path and stays a cuDNN network; this only removes glue.
Reference module: src/ops/e4e/encoders/psp_encoders.py:125-216 (eval mode, e4e_arch.py:256-258).
"""
import copy

import torch
from torch import nn


def _fold(conv, bn):
    """conv (no activation in between) followed by eval BatchNorm2d -> one conv with bias."""
    w = conv.weight.detach().float()
    b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
    g = bn.weight.detach().float() * torch.rsqrt(bn.running_var.detach().float() + bn.eps)
    fused = nn.Conv2d(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding, bias=True)
    fused.weight = nn.Parameter(w * g.reshape(-1, 1, 1, 1), requires_grad=False)
    fused.bias = nn.Parameter((b - bn.running_mean.detach().float()) * g + bn.bias.detach().float(), requires_grad=False)
    return fused


def _fold_sequential(seq):
    mods = list(seq.children())
    out, i = [], 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, nn.Conv2d) and i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm2d):
            out.append(_fold(m, mods[i + 1]))
            i += 2
        else:
            out.append(m)
            i += 1
    return nn.Sequential(*out)


def state_key(module):
    """Cache key of a module's state: versions AND storage addresses (`p.data = t`, how checkpoints are often loaded, keeps the version)."""
    return tuple((p._version, p.data_ptr()) for p in module.parameters()) + tuple((b._version, b.data_ptr()) for b in module.buffers())


@torch.no_grad()
def build(encoder, dtype=torch.bfloat16):
    enc = copy.deepcopy(encoder).eval()
    enc.input_layer = _fold_sequential(enc.input_layer)
    for blk in enc.body:
        blk.res_layer = _fold_sequential(blk.res_layer)
        if isinstance(blk.shortcut_layer, nn.Sequential):
            blk.shortcut_layer = _fold_sequential(blk.shortcut_layer)
    enc = enc.to(dtype).to(memory_format=torch.channels_last)
    for p in enc.parameters():
        p.requires_grad_(False)
    return enc
