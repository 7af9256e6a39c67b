"""Oracle (test infrastructure): the full OOD inversion forward, functional.

Restates src/archs/OOD_faceGAN_e4e_arch.py:224-347 (forward, callback, blending_mask, blend) and
src/ops/e4e/encoders/psp_encoders.py:34-56,125-216 (Encoder4Editing, GradualStyleBlock) with
bottleneck_IR_SE / SEModule of src/ops/e4e/encoders/helpers.py:59-76,476-501.
"""
import math

import torch
import torch.nn.functional as F

from . import samm
from .stylegan import channel_map, equal_linear, generator_forward, synthetic_generator_state

IR50_UNITS = [(64, 64, 2)] + [(64, 64, 1)] * 2 + [(64, 128, 2)] + [(128, 128, 1)] * 3 + \
             [(128, 256, 2)] + [(256, 256, 1)] * 13 + [(256, 512, 2)] + [(512, 512, 1)] * 2


def _bn(sd, p, x, eps=1e-5):
    return F.batch_norm(x, sd[f'{p}running_mean'], sd[f'{p}running_var'], sd[f'{p}weight'],
                        sd[f'{p}bias'], False, 0.0, eps)


def _ir_se_unit(sd, p, x, cin, depth, stride):
    """bottleneck_IR_SE, BatchNorm in eval mode.  e4e helpers.py:476-501, 59-76."""
    if cin == depth:
        sc = x[:, :, ::stride, ::stride]          # MaxPool2d(kernel 1, stride)
    else:
        sc = _bn(sd, f'{p}shortcut_layer.1.', F.conv2d(x, sd[f'{p}shortcut_layer.0.weight'], stride=stride))
    r = _bn(sd, f'{p}res_layer.0.', x)
    r = F.conv2d(r, sd[f'{p}res_layer.1.weight'], padding=1)
    r = F.prelu(r, sd[f'{p}res_layer.2.weight'])
    r = F.conv2d(r, sd[f'{p}res_layer.3.weight'], stride=stride, padding=1)
    r = _bn(sd, f'{p}res_layer.4.', r)
    g = r.mean(dim=(2, 3), keepdim=True)
    g = torch.sigmoid(F.conv2d(F.relu(F.conv2d(g, sd[f'{p}res_layer.5.fc1.weight'])),
                               sd[f'{p}res_layer.5.fc2.weight']))
    return r * g + sc


def _style_head(sd, p, x):
    """GradualStyleBlock: stride-2 convs + LeakyReLU(0.01) down to 1x1, then EqualLinear.  :34-56."""
    n = int(math.log2(x.shape[-1]))
    for j in range(n):
        x = F.leaky_relu(F.conv2d(x, sd[f'{p}convs.{2 * j}.weight'], sd[f'{p}convs.{2 * j}.bias'],
                                  stride=2, padding=1), 0.01)
    return equal_linear(x.reshape(-1, x.shape[1]), sd[f'{p}linear.weight'], sd[f'{p}linear.bias'])


def _up_add(x, y):
    """e4e helpers.py:504-521."""
    return F.interpolate(x, size=y.shape[-2:], mode='bicubic', align_corners=True) + y


def e4e_encoder(sd, x, prefix='encoder.', n_styles=18):
    """Encoder4Editing.forward(return_feats=True) at ProgressiveStage.Inference.  :178-216."""
    p = prefix
    x = F.conv2d(x, sd[f'{p}input_layer.0.weight'], padding=1)
    x = F.prelu(_bn(sd, f'{p}input_layer.1.', x), sd[f'{p}input_layer.2.weight'])
    feats = [x]
    taps = {}
    for i, (cin, depth, stride) in enumerate(IR50_UNITS):
        x = _ir_se_unit(sd, f'{p}body.{i}.', x, cin, depth, stride)
        if i in (2, 6, 20, 23):
            feats.append(x)
            taps[i] = x
    c1, c2, c3 = taps[6], taps[20], taps[23]
    w0 = _style_head(sd, f'{p}styles.0.', c3)
    w = [w0.clone() for _ in range(n_styles)]
    f = c3
    for i in range(1, n_styles):
        if i == 3:
            f = _up_add(c3, F.conv2d(c2, sd[f'{p}latlayer1.weight'], sd[f'{p}latlayer1.bias']))
            p2 = f
        elif i == 7:
            f = _up_add(p2, F.conv2d(c1, sd[f'{p}latlayer2.weight'], sd[f'{p}latlayer2.bias']))
        w[i] = w[i] + _style_head(sd, f'{p}styles.{i}.', f)
    return torch.stack(w, dim=1), feats


def ood_forward(sd, x, size=1024, warp_scale=0.08, cycle_align=2, mod_size=256, truncation=1.0,
                strict_rng=True, latents=None, enc_feats=None):
    """ood_faceGAN_e4e.forward (E4E encoder, modulation_type NOISE, blend_with_gen, one blend).
    src/archs/OOD_faceGAN_e4e_arch.py:245-313.  Returns (out, lats, aligns dict).

    strict_rng reproduces the reference's wasted `randn_like(image)` draw inside the callback
    (e4e_arch.py:234) so that seeded CPU runs consume the RNG stream identically.
    `latents` / `enc_feats` bypass the encoder (for generator+SAMM-only parity cases).
    """
    if latents is None:
        w, feats = e4e_encoder(sd, F.interpolate(x, (256, 256), mode='bilinear'))
    else:
        w, feats = latents, enc_feats
    lats = w + sd['avg_latent'].reshape(1, 1, -1) + sd['delta_latent']
    if truncation < 1.0:
        lats = sd['avg_latent'].reshape(1, 1, -1) * (1.0 - truncation) + lats * truncation
    enc = [F.conv2d(feats[i], sd[f'feats_conv.{i}.weight'], sd[f'feats_conv.{i}.bias']) for i in range(4)]
    # feats2condition: number of conditioned levels (e4e_arch.py:214-222)
    n_cond = min(max(1 + int(math.log2(mod_size)) - int(math.log2(enc[-1].shape[-1])), 0), 4) if mod_size > 0 else 0
    cond_layers = [2 * (k + 2) + 1 for k in range(n_cond)]
    aligns = {}

    def hook(ci, image, noise, noise_weight, style):
        ind = ci + 1
        if strict_rng:
            torch.randn_like(image)
        coarse = aligns[ind - 1] if ind > 1 else None
        aligned, field = samm.spm_warp(sd, f'modulation.{4 - ind}.alignment.', enc[-ind], image, coarse,
                                       warp_scale, cycle_align)
        aligns[ind] = field
        return (aligned - image + noise * noise_weight) / noise_weight

    gen = generator_forward(sd, lats, size, cond_layers=cond_layers, hook=hook, prefix='generator.')
    if n_cond > 0:
        alpha = samm.compose_masks([aligns[k] for k in sorted(aligns)], size)
        aligns[1024] = alpha.repeat(1, 3, 1, 1)
        out = samm.blend(alpha, x, gen)
    else:
        out = gen
    return out, lats, aligns


def synthetic_ood_state(size=1024, seed=0, with_encoder=True):
    """Random-init weights under the reference's key names.  Generator as in
    synthetic_generator_state; SAMM convs xavier-normal (SAMM/helpers.py:124-127), norms
    weight=1/bias=0 perturbed by 0.1*N(0,1); encoder convs kaiming-like so activations stay O(1).
    """
    g = torch.Generator().manual_seed(seed + 1000)

    def rn(*s):
        return torch.randn(*s, generator=g)

    sd = {f'generator.{k}': v for k, v in synthetic_generator_state(size, seed=seed).items()}
    sd['avg_latent'] = 0.1 * rn(1, 512)
    sd['delta_latent'] = torch.zeros(1, 18, 512)
    ch = channel_map(2)
    enc_ch = [64, 64, 128, 256]
    for i, r in enumerate([256, 128, 64, 32]):
        sd[f'feats_conv.{i}.weight'] = rn(ch[r], enc_ch[i], 1, 1) / math.sqrt(enc_ch[i])
        sd[f'feats_conv.{i}.bias'] = 0.1 * rn(ch[r])

    def xavier(co, ci, k):
        return rn(co, ci, k, k) * math.sqrt(2.0 / ((ci + co) * k * k))

    def norm(p, c):
        sd[f'{p}weight'] = 1.0 + 0.1 * rn(c)
        sd[f'{p}bias'] = 0.1 * rn(c)

    for m, r in enumerate([256, 128, 64, 32]):
        c2 = 2 * ch[r]
        p = f'modulation.{m}.'
        sd[f'{p}weight'] = torch.ones(1)
        sd[f'{p}noiseInj.weight'] = torch.zeros(1)
        sd[f'{p}alignment.blur.kernel'] = torch.outer(torch.tensor([1., 3., 3., 1.]), torch.tensor([1., 3., 3., 1.])) / 64.0
        b0, b1 = f'{p}alignment.body.body.0.', f'{p}alignment.body.body.1.'
        norm(f'{b0}res_layer.0.', c2)
        sd[f'{b0}res_layer.1.weight'] = xavier(c2, c2, 3)
        sd[f'{b0}res_layer.2.weight'] = torch.full((c2,), 0.25)
        sd[f'{b0}res_layer.3.weight'] = xavier(c2, c2, 3)
        norm(f'{b0}res_layer.4.', c2)
        sd[f'{b1}shortcut_layer.0.weight'] = xavier(3, c2, 1)
        norm(f'{b1}shortcut_layer.1.', 3)
        norm(f'{b1}res_layer.0.', c2)
        sd[f'{b1}res_layer.1.weight'] = xavier(3, c2, 3)
        sd[f'{b1}res_layer.2.weight'] = torch.full((3,), 0.25)
        sd[f'{b1}res_layer.3.weight'] = xavier(3, 3, 3)
        norm(f'{b1}res_layer.4.', 3)
    if with_encoder:
        def bn(p, c):
            sd[f'{p}weight'] = 1.0 + 0.1 * rn(c)
            sd[f'{p}bias'] = 0.1 * rn(c)
            sd[f'{p}running_mean'] = 0.1 * rn(c)
            sd[f'{p}running_var'] = 1.0 + 0.1 * torch.rand(c, generator=g)
            sd[f'{p}num_batches_tracked'] = torch.zeros((), dtype=torch.long)

        def kconv(co, ci, k):
            return rn(co, ci, k, k) / math.sqrt(ci * k * k)

        e = 'encoder.'
        sd[f'{e}input_layer.0.weight'] = kconv(64, 3, 3)
        bn(f'{e}input_layer.1.', 64)
        sd[f'{e}input_layer.2.weight'] = torch.full((64,), 0.25)
        for i, (cin, depth, stride) in enumerate(IR50_UNITS):
            p = f'{e}body.{i}.'
            if cin != depth:
                sd[f'{p}shortcut_layer.0.weight'] = kconv(depth, cin, 1)
                bn(f'{p}shortcut_layer.1.', depth)
            bn(f'{p}res_layer.0.', cin)
            sd[f'{p}res_layer.1.weight'] = kconv(depth, cin, 3)
            sd[f'{p}res_layer.2.weight'] = torch.full((depth,), 0.25)
            sd[f'{p}res_layer.3.weight'] = kconv(depth, depth, 3)
            bn(f'{p}res_layer.4.', depth)
            sd[f'{p}res_layer.5.fc1.weight'] = kconv(depth // 16, depth, 1)
            sd[f'{p}res_layer.5.fc2.weight'] = kconv(depth, depth // 16, 1)
        for i in range(18):
            spatial = 16 if i < 3 else (32 if i < 7 else 64)
            p = f'{e}styles.{i}.'
            for j in range(int(math.log2(spatial))):
                sd[f'{p}convs.{2 * j}.weight'] = kconv(512, 512, 3)
                sd[f'{p}convs.{2 * j}.bias'] = 0.1 * rn(512)
            sd[f'{p}linear.weight'] = rn(512, 512)
            sd[f'{p}linear.bias'] = torch.zeros(512)
        sd[f'{e}latlayer1.weight'] = kconv(512, 256, 1)
        sd[f'{e}latlayer1.bias'] = 0.1 * rn(512)
        sd[f'{e}latlayer2.weight'] = kconv(512, 128, 1)
        sd[f'{e}latlayer2.bias'] = 0.1 * rn(512)
    return sd
