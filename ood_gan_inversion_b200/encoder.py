"""E4E encoder (IR-SE50 trunk + FPN + 18 GradualStyleBlocks) with the reference's module tree and state-dict keys.

Reference: src/ops/e4e/encoders/psp_encoders.py:34-56,125-216 and helpers.py:26-76,476-501.  The encoder sits IN FRONT
of the hot path and stays a PyTorch/cuDNN network; it is here so that the full
inversion pipeline runs without the reference on the path.  In bf16 mode it runs under autocast in channels_last.
"""
import math
from collections import namedtuple
from enum import Enum

import torch
from torch import nn
from torch.nn import functional as F

from . import kernels as K
from .samm import BN
from .stylegan import EqualLinear


class ProgressiveStage(Enum):
    """psp_encoders.py:13-32"""
    WTraining = 0
    Delta1Training = 1
    Delta2Training = 2
    Delta3Training = 3
    Delta4Training = 4
    Delta5Training = 5
    Delta6Training = 6
    Delta7Training = 7
    Delta8Training = 8
    Delta9Training = 9
    Delta10Training = 10
    Delta11Training = 11
    Delta12Training = 12
    Delta13Training = 13
    Delta14Training = 14
    Delta15Training = 15
    Delta16Training = 16
    Delta17Training = 17
    Inference = 18


Bottleneck = namedtuple('Block', ['in_channel', 'depth', 'stride'])


def get_block(in_channel, depth, num_units, stride=2):
    return [Bottleneck(in_channel, depth, stride)] + [Bottleneck(depth, depth, 1) for _ in range(num_units - 1)]


def get_blocks(num_layers):
    """helpers.py:34-56"""
    units = {50: (3, 4, 14, 3), 100: (3, 13, 30, 3), 152: (3, 8, 36, 3)}
    if num_layers not in units:
        raise ValueError(f'Invalid number of layers: {num_layers}. Must be one of [50, 100, 152]')
    n = units[num_layers]
    return [get_block(64, 64, n[0]), get_block(64, 128, n[1]), get_block(128, 256, n[2]), get_block(256, 512, n[3])]


class SEModule(nn.Module):
    """helpers.py:59-76"""

    def __init__(self, channels, reduction):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc1 = nn.Conv2d(channels, channels // reduction, kernel_size=1, padding=0, bias=False)
        self.relu = nn.ReLU(inplace=True)
        self.fc2 = nn.Conv2d(channels // reduction, channels, kernel_size=1, padding=0, bias=False)
        self.sigmoid = nn.Sigmoid()

    def forward(self, x):
        return x * self.sigmoid(self.fc2(self.relu(self.fc1(self.avg_pool(x)))))


class bottleneck_IR_SE(nn.Module):
    """helpers.py:476-501"""

    def __init__(self, in_channel, depth, stride, bn=True):
        super().__init__()
        if in_channel == depth:
            self.shortcut_layer = nn.MaxPool2d(1, stride)
        else:
            self.shortcut_layer = nn.Sequential(nn.Conv2d(in_channel, depth, (1, 1), stride, bias=False), BN(depth, bn=bn))
        self.res_layer = nn.Sequential(BN(in_channel, bn=bn), nn.Conv2d(in_channel, depth, (3, 3), (1, 1), 1, bias=False),
                                       nn.PReLU(depth), nn.Conv2d(depth, depth, (3, 3), stride, 1, bias=False),
                                       BN(depth, bn=bn), SEModule(depth, 16))

    def forward(self, x):
        if isinstance(self.shortcut_layer, nn.MaxPool2d):       # MaxPool2d(1, s) == strided subsampling: no kernel needed
            s = self.shortcut_layer.stride
            return self.res_layer(x) + (x if s == 1 else x[:, :, ::s, ::s])
        return self.res_layer(x) + self.shortcut_layer(x)


class GradualStyleBlock(nn.Module):
    """psp_encoders.py:34-56"""

    def __init__(self, in_c, out_c, spatial):
        super().__init__()
        self.out_c = out_c
        self.spatial = spatial
        num_pools = int(math.log2(spatial))
        modules = [nn.Conv2d(in_c, out_c, kernel_size=3, stride=2, padding=1), nn.LeakyReLU()]
        for _ in range(num_pools - 1):
            modules += [nn.Conv2d(out_c, out_c, kernel_size=3, stride=2, padding=1), nn.LeakyReLU()]
        self.convs = nn.Sequential(*modules)
        self.linear = EqualLinear(out_c, out_c, lr_mul=1)

    def forward(self, x):
        x = self.convs(x).reshape(-1, self.out_c)
        with torch.autocast('cuda', enabled=False):        # the W+ latents are always produced in fp32
            return F.linear(x.float(), self.linear.weight.float() * self.linear.scale, self.linear.bias.float() * self.linear.lr_mul)


def _upsample_add(x, y):
    """helpers.py:504-521: bicubic (align_corners=True) upsample of the top feature map + lateral map; fused NHWC kernel
    (ATen's bicubic on channels_last tensors takes 40 ms per call at batch 16)."""
    if not x.is_cuda:
        raise RuntimeError('ood_gan_inversion_b200 is CUDA-only')
    dt = y.dtype if y.dtype in (torch.float32, torch.bfloat16) else torch.float32
    xn = x.to(dt).permute(0, 2, 3, 1).contiguous()
    yn = y.to(dt).permute(0, 2, 3, 1).contiguous()
    return K.bicubic_up_add(xn, yn).permute(0, 3, 1, 2)          # channels_last view of the NHWC result


class Encoder4Editing(nn.Module):
    """psp_encoders.py:125-216"""

    def __init__(self, num_layers, mode='ir', opts=None, bn=True):
        super().__init__()
        assert num_layers in [50, 100, 152] and mode in ['ir', 'ir_se']
        if mode != 'ir_se':
            raise NotImplementedError("ood_gan_inversion_b200: the OOD arch builds the encoder with mode='ir_se' only")
        blocks = get_blocks(num_layers)
        self.input_layer = nn.Sequential(nn.Conv2d(3, 64, (3, 3), 1, 1, bias=False), BN(64, bn=bn), nn.PReLU(64))
        self.channels = [64]
        modules = []
        for block in blocks:
            for bt in block:
                modules.append(bottleneck_IR_SE(bt.in_channel, bt.depth, bt.stride, bn=bn))
            self.channels.append(block[-1].depth)
        self.body = nn.Sequential(*modules)
        self.styles = nn.ModuleList()
        stylegan_size = opts['stylegan_size'] if isinstance(opts, dict) else opts.stylegan_size
        self.style_count = 2 * int(math.log(stylegan_size, 2)) - 2
        self.coarse_ind, self.middle_ind = 3, 7
        for i in range(self.style_count):
            self.styles.append(GradualStyleBlock(512, 512, 16 if i < self.coarse_ind else (32 if i < self.middle_ind else 64)))
        self.latlayer1 = nn.Conv2d(256, 512, kernel_size=1, stride=1, padding=0)
        self.latlayer2 = nn.Conv2d(128, 512, kernel_size=1, stride=1, padding=0)
        self.progressive_stage = ProgressiveStage.Inference

    def get_deltas_starting_dimensions(self):
        return list(range(self.style_count))

    def set_progressive_stage(self, new_stage):
        self.progressive_stage = new_stage

    def forward(self, x, **kwargs):
        x = self.input_layer(x)
        feats = [x]
        for i, layer in enumerate(self.body):
            x = layer(x)
            if i == 2:
                feats.append(x)
            elif i == 6:
                c1 = x
                feats.append(x)
            elif i == 20:
                c2 = x
                feats.append(x)
            elif i == 23:
                c3 = x
                feats.append(x)
        w0 = self.styles[0](c3)
        w = [w0] * self.style_count
        stage = self.progressive_stage.value
        features = c3
        for i in range(1, min(stage + 1, self.style_count)):
            if i == self.coarse_ind:
                p2 = _upsample_add(c3, self.latlayer1(c2))
                features = p2
            elif i == self.middle_ind:
                p1 = _upsample_add(p2, self.latlayer2(c1))
                features = p1
            w[i] = w0 + self.styles[i](features)
        w = torch.stack(w, dim=1)
        if kwargs.get('return_feats', False):
            return w, feats
        return w
