"""Drop-in for the reference's `ood_faceGAN_e4e` (src/archs/OOD_faceGAN_e4e_arch.py:27-347).

Same constructor arguments, `forward(x, **kwargs) -> (out, lats)`, `self.aligns` / `self.feats` / `self.lats`
side outputs and state-dict keys (encoder.*, feats_conv.*, modulation.*, generator.*, avg_latent, delta_latent).
The glue is re-expressed for throughput (SURVEY finding 13): no `torch.cuda.empty_cache()` storms, the alignment
callback stays in NHWC (`aligned_nhwc`), and the four-level mask compose + clip + ID/OOD blend is one kernel.
"""
import math

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from . import kernels as K
from . import stylegan as sg
from .encoder import Encoder4Editing, ProgressiveStage
from .samm import StyledscaleNshfitBlock
from .stylegan import Generator


class _AlignCallback:
    """The reference's feats2condition_callback (e4e_arch.py:224-242) as an object.

    __call__ keeps the reference contract (NCHW image in, `(aligned - image + noise*w)/w` out) so a foreign generator
    can use it; `aligned_nhwc` is the fused route used by this package's Generator: it returns the aligned features
    and the noise injection stays in the following kernel (net effect image := aligned + w*noise)."""

    def __init__(self, owner):
        self.owner = owner

    def _level(self, index):
        ind = index + 1
        o = self.owner
        return ind, o.feats[-ind], o.modulation[-ind], (o.aligns[ind - 1] if ind > 1 else None)

    def aligned_nhwc(self, image_nhwc, **kwargs):
        ind, feat, mod, coarse = self._level(kwargs.get('index'))
        if self.owner.strict_rng:       # the reference evaluates randn_like(image) eagerly (e4e_arch.py:234)
            b, h, w, c = image_nhwc.shape
            torch.empty(b, c, h, w, device=image_nhwc.device, dtype=torch.float32).normal_()
        aligned, field = mod.forward_nhwc(feat, image_nhwc, coarse)      # differentiable when image / coarse field / feat carry a graph
        self.owner.aligns[ind] = field
        return aligned

    def __call__(self, image, **kwargs):
        ind, feat, mod, coarse = self._level(kwargs.get('index'))
        noise = kwargs.get('noise', torch.randn_like(image))
        noise_weight = kwargs.get('noise_weight', 1)
        condition, align = mod(feat, kwargs.get('style'), image=image, aligned=coarse)
        self.owner.aligns[ind] = align
        return (condition - image + noise * noise_weight) / noise_weight


class ood_faceGAN_e4e(nn.Module):
    def __init__(self,
                 # generator opts
                 out_size=1024, style_dim=512, n_mlp=8, channel_multiplier=2, narrow=1, merge='',
                 StyleGAN_pth=None, StyleGAN_pth_key='params_ema',
                 # augmentation
                 aug_alignment=False, aug_inputcolor=False,
                 # encoder opts
                 stage='Inference', encoder='E4E', E4E_pth=None, avg_latent_pth=None,
                 optim_delta_latent=False, delta_latent_pth=None,
                 # modulation opts
                 enable_modulation=True, modulation_type='NOISE', warp_scale=0.02,
                 blend_with_gen=True, ModSize=None,
                 # training opts
                 progressiveModSize=[16, 32, 64, 128, 256], progressiveStart=20000,
                 progressiveStep=2000, progressiveStageSteps=[999999999], eval_path_length=None,
                 **kwargs):
        super().__init__()
        if encoder != 'E4E':
            raise NotImplementedError("ood_gan_inversion_b200: only encoder='E4E' is built (restyle / FeatureStyle archs are out "
                                      'of scope, SURVEY section 2.1 row 9)')
        if modulation_type != 'NOISE':
            raise NotImplementedError("ood_gan_inversion_b200: only modulation_type='NOISE' is implemented")
        if aug_alignment or aug_inputcolor:
            raise NotImplementedError('ood_gan_inversion_b200: training-time augmentations are out of scope')
        self.encoder_type = encoder
        log_outsize = int(math.log(out_size, 2))
        self.style_cnt = log_outsize * 2 - 2
        self.style_dim = style_dim
        cm = channel_multiplier
        self.channels = {4: int(512 * narrow), 8: int(512 * narrow), 16: int(512 * narrow), 32: int(512 * narrow),
                         64: int(256 * cm * narrow), 128: int(128 * cm * narrow), 256: int(64 * cm * narrow),
                         512: int(32 * cm * narrow), 1024: int(16 * cm * narrow), 2048: int(8 * cm * narrow)}
        self.encoder = Encoder4Editing(num_layers=50, mode='ir_se', opts={'stylegan_size': out_size}, bn=True)
        if enable_modulation:
            self.feats_conv = nn.ModuleList()
            featsize = 256
            for i in range(4):
                self.feats_conv.append(nn.Conv2d(self.encoder.channels[i], self.channels[featsize], kernel_size=1))
                featsize //= 2
        self.aligns = {}
        self.log_outsize = int(math.log(256, 2))
        self.randomTransform = None
        self.colorTransform = None
        if enable_modulation:
            self.modulation = nn.ModuleList()
            self.progressiveModSize = list(progressiveModSize)
            self.modulation_type = modulation_type
            self.blend_with_gen = blend_with_gen
            self.blend_cnt = kwargs.get('blend_cnt', 1)
            self.skip_SA = kwargs.get('skip_SA', False)
            self.ModSize = self.progressiveModSize.pop(0) if ModSize is None else ModSize
            for i in range(self.log_outsize, 4, -1):
                chn = self.channels[2 ** i]
                self.modulation.append(StyledscaleNshfitBlock(chn, chn, style_dim, scale=warp_scale,
                                                              btn=kwargs.get('mod_btn', None),
                                                              cycle_align=kwargs.get('cycle_align', 1),
                                                              diff_fAndg=kwargs.get('diff_fAndg', True)))
        else:
            self.modulation = None
            self.ModSize = 0
        self.generator = Generator(size=out_size, n_mlp=n_mlp, style_dim=style_dim, channel_multiplier=channel_multiplier)
        self.avg_latent = nn.Parameter(torch.zeros((1, style_dim)), requires_grad=False)
        if optim_delta_latent:
            self.delta_latent = nn.Parameter(torch.randn((1, 18, style_dim)) * 0.1, requires_grad=True)
        else:
            self.delta_latent = nn.Parameter(torch.zeros((1, 18, style_dim)), requires_grad=False)
        self.encoder.progressive_stage = ProgressiveStage[stage]
        self.progressiveStageSteps = progressiveStageSteps
        if self.progressiveStageSteps is None:
            self.progressiveStageSteps = [progressiveStart + progressiveStep * i for i in range(self.style_cnt)]
        if StyleGAN_pth is not None:
            self.generator.load_state_dict(torch.load(StyleGAN_pth, map_location='cpu')[StyleGAN_pth_key], strict=False)
        if E4E_pth is not None:
            enc = torch.load(E4E_pth, map_location='cpu')['state_dict']
            self.encoder.load_state_dict({k[len('encoder.'):]: v for k, v in enc.items() if 'encoder.' in k}, strict=True)
        if avg_latent_pth is not None:
            self.avg_latent.data = torch.load(avg_latent_pth, map_location='cpu')
        if delta_latent_pth is not None:
            self.delta_latent.data = torch.load(delta_latent_pth, map_location='cpu')
        # e4e_arch.py:146-151: at the Inference stage with modulation the reference marks the encoder's codes as requiring grad,
        # so that a forward outside torch.no_grad() builds the graph of the whole pipeline w.r.t. them (path-length evaluation)
        if eval_path_length is not None:
            self.eval_path_length = bool(eval_path_length)
        else:
            self.eval_path_length = bool(enable_modulation) and self.encoder.progressive_stage == ProgressiveStage.Inference
        self.strict_rng = False          # True: replay the reference's RNG stream bit-for-bit (wasted draws included)
        self._callback = _AlignCallback(self)
        self.feats, self.lats, self.ori_lats = None, None, None

    # ---- reference helpers kept for API parity -------------------------------------------------------------------
    def update_stage(self, step, logger=None):
        """e4e_arch.py:155-182"""
        while len(self.progressiveStageSteps) > 0 and step > self.progressiveStageSteps[0]:
            self.progressiveStageSteps.pop(0)
            if self.encoder.progressive_stage.value < self.style_cnt:
                self.encoder.progressive_stage = ProgressiveStage(self.encoder.progressive_stage.value + 1)
            if self.modulation is not None and len(self.progressiveModSize) > 0 and self.ModSize < self.progressiveModSize[0]:
                self.ModSize = self.progressiveModSize.pop(0)

    def get_style_mlp(self, x):
        return self.generator.style(x)

    def random_gen(self, batch_size=1, gen=True):
        with torch.no_grad():
            style = torch.randn((batch_size, self.style_dim), device=self.avg_latent.device)
            lats = self.get_style_mlp(style).unsqueeze(1).repeat(1, self.style_cnt, 1)
            out = self.generator(lats, input_is_tensor=True, input_is_latent=True)[0] if gen else None
        return out, lats

    def feats2condition(self, feats, **kwargs):
        """e4e_arch.py:214-222"""
        conditions = []
        if self.ModSize > 0:
            max_size = int(np.floor(math.log(self.ModSize, 2)))
            min_size = int(np.floor(math.log(feats[-1].shape[-1], 2)))
            for _ in range(min(max((1 + max_size - min_size), 0), len(feats))):
                conditions.append([None, None])
        return conditions

    def feats2condition_callback(self, image, **kwargs):
        return self._callback(image, **kwargs)

    # ---- forward -------------------------------------------------------------------------------------------------
    def encode(self, x, with_offsets=False, fold_offsets=True):
        """E4E encoder on the bilinear 256x256 thumbnail -> (w [B,18,512] fp32, feats).  e4e_arch.py:256-258.
        bf16 mode runs a cached inference copy on this library's kernels (encoder_fast.FastEncoder).
        with_offsets=True returns (w + avg_latent + delta_latent, feats, True) when the offsets (e4e_arch.py:261) could be added
        inside the W+ assembly kernel (bf16 mode, nothing to differentiate), else (w, feats, False)."""
        bf16 = sg.get_precision() == 'bf16'
        fold = bool(with_offsets) and bool(fold_offsets) and bf16 and \
            not (torch.is_grad_enabled() and (self.delta_latent.requires_grad or self.avg_latent.requires_grad))
        with torch.no_grad():
            self.encoder.eval()
            if bf16:
                from . import encoder_fast, encoder_infer
                key = (x.device, encoder_infer.state_key(self.encoder))
                if getattr(self, '_enc_infer', None) is None or self._enc_infer[0] != key:
                    object.__setattr__(self, '_enc_infer', (key, encoder_fast.FastEncoder(self.encoder)))
                enc = self._enc_infer[1]
                enc.progressive_stage = self.encoder.progressive_stage
                # bilinear 1024 -> 256 thumbnail, NHWC, bf16, channels padded to the first convolution's K granule: one kernel
                w, feats = enc(None, return_feats=True, thumb=K.thumbnail_nhwc(x.detach(), dtype=encoder_fast.ENC_DT),
                               avg=self.avg_latent.detach() if fold else None, delta=self.delta_latent.detach()[0] if fold else None)
            else:
                small = F.interpolate(x, (256, 256), mode='bilinear')
                with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                    w, feats = self.encoder(small, return_feats=True)
        if with_offsets:
            return w.float(), feats, fold
        return w.float(), feats

    def _feats_conv_nhwc(self, i, feat):
        """feats_conv[i] (1x1 convolution + bias, e4e_arch.py:109-113) on the tcgen05 1x1 form; NCHW view of an NHWC result."""
        conv = self.feats_conv[i]
        dt = feat.dtype if feat.dtype in (torch.bfloat16, torch.float16) else torch.bfloat16      # the encoder's storage type
        key = (conv.weight.device, conv.weight._version, conv.weight.data_ptr(), conv.bias._version, conv.bias.data_ptr(), dt)
        cache = self.__dict__.setdefault('_fc_cache', {})
        if cache.get(i, (None,))[0] != key:
            with torch.no_grad():
                cache[i] = (key, K.pack_conv1x1_weight(conv.weight.detach(), dt, False), conv.bias.detach().float().contiguous())
        _, w, b = cache[i]
        with torch.no_grad():
            x = feat.to(dt).permute(0, 2, 3, 1).contiguous()          # NCHW view of the encoder's NHWC tensor: no copy
            y, _ = K.conv3x3(x, w, conv.out_channels, transposed=4, bias=b, tag='encoder_conv', out_dtype=torch.bfloat16)
        return y.permute(0, 3, 1, 2)

    def _feats_conv_diff(self, i, feat):
        """feats_conv[i] on the gradient path (the reference trains it together with `modulation`: options/train/E4E_Face.yml:123-125
        fixes only the generator and the encoder): 1x1 convolution + bias as diff_ops Functions, so that its weight / bias gradients
        come from ood_conv_wgrad and the bias reduction.  The encoder features themselves are constants (e4e_arch.py:256-258)."""
        from . import diff_ops as D
        conv = self.feats_conv[i]
        x = feat.detach().permute(0, 2, 3, 1).contiguous().to(sg._act_dtype())
        return D.bias_add(D.conv(x, conv.weight, '1x1'), conv.bias).permute(0, 3, 1, 2)

    def forward(self, x, **kwargs):
        if kwargs.get('random_gen', False):
            return self.random_gen(batch_size=kwargs.get('batch_size', 1), gen=kwargs.get('gen', True))
        step = kwargs.get('step', None)
        if step is not None:
            self.update_stage(step, kwargs.get('logger', None))
        bf16 = sg.get_precision() == 'bf16'
        grad_route = torch.is_grad_enabled() and (self.eval_path_length or self.delta_latent.requires_grad)
        lats, feats, offsets_in = self.encode(x, with_offsets=True, fold_offsets=not grad_route)
        if grad_route and self.eval_path_length:
            lats.requires_grad = True                   # e4e_arch.py:258-259
        if not offsets_in:
            lats = lats + self.avg_latent.reshape(1, 1, -1) + self.delta_latent
        truncation = kwargs.get('truncation', 1.0)
        if truncation < 1.0:
            lats = self.avg_latent.reshape(1, 1, -1) * (1. - truncation) + (lats * truncation)
        self.ori_lats = lats
        if self.modulation is None:
            out, _ = self.generator(lats, input_is_tensor=True, input_is_latent=True)
            return out, lats
        train_feats = torch.is_grad_enabled() and any(p.requires_grad for p in self.feats_conv.parameters())
        if train_feats:
            self.feats = [self._feats_conv_diff(i, feats[i]) for i in range(4)]
        elif bf16:
            self.feats = [self._feats_conv_nhwc(i, feats[i]) for i in range(4)]
        else:
            with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                self.feats = [self.feats_conv[i](feats[i]) for i in range(4)]
        self.lats = lats
        self.aligns = {}
        conditions = self.feats2condition(self.feats)
        cond_ind = [(2 * (k + 2)) + 1 for k in range(len(conditions))]
        out, _ = self.generator(lats, input_is_tensor=True, input_is_latent=True, conditions=conditions,
                                cond_layers=cond_ind, cond_type=self.modulation_type, callback=self._callback,
                                align_aug=False)
        if self.blend_with_gen:
            if self.skip_SA:
                out, _ = self.generator(lats, input_is_tensor=True, input_is_latent=True)
            keys = sorted(self.aligns.keys())
            if keys:
                fields = [self.aligns[k] for k in keys]
                for _ in range(self.blend_cnt):
                    if torch.is_grad_enabled() and (out.requires_grad or any(f.requires_grad for f in fields)):
                        from . import samm_grad
                        out, alpha = samm_grad.mask_blend(fields, x.detach(), out)      # backward: ood_mask_blend_bwd
                    else:
                        out, alpha = K.mask_blend(fields, x.detach(), out)
                self.aligns[1024] = alpha.expand(-1, 3, -1, -1)    # 3-channel view, not materialised
        return out, lats

    def blending_mask(self):
        """e4e_arch.py:315-339 (standalone form; forward() fuses it with the blend)."""
        keys = sorted(k for k in self.aligns.keys() if k != 1024)
        if not keys:
            return None
        f0 = self.aligns[keys[0]]
        size = self.generator.size
        zeros = torch.zeros(f0.shape[0], 3, size, size, device=f0.device)
        _, alpha = K.mask_blend([self.aligns[k] for k in keys], zeros, zeros)
        self.aligns[1024] = alpha.expand(-1, 3, -1, -1)
        return alpha

    def blend(self, target, output, detach=True, alpha_scale=None):
        """e4e_arch.py:341-347"""
        if alpha_scale is None:
            return None
        return alpha_scale * target + output * (1 - alpha_scale)
