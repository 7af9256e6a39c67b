"""Differentiable forms of the SAMM gather / blend kernels ("grid_sample grads").

    field = field_step(z, prev, coarse, scale)          # SAMM/helpers.py:62-77,149-166: heads + FIR + accumulate / clip / PRM
    aligned = warp_mix(gen_nhwc, field)                 # SAMM/helpers.py:168-177: grid_sample + alpha mix
    out, alpha = mask_blend(fields, x, gen)             # OOD_faceGAN_e4e_arch.py:315-347: mask pyramid, clip, blend

Forward = the kernels of the inference path (ood_field_step, ood_warp_mix, ood_mask_blend); backward = ood_field_step_bwd /
ood_warp_mix_bwd / ood_mask_blend_bwd,
whose per-item bodies are checked against torch.autograd of the reference arithmetic on the CPU (tests/test_samm_bwd_cpu.py)
and through the C ABI on the GPU (tests/test_zz_samm_bwd_gpu.py).  The inference modules (samm.py, arch.py) keep calling the plain kernels.
"""
import torch
from torch.autograd import Function

from . import kernels as K


class _WarpMix(Function):
    @staticmethod
    def forward(ctx, gen, field):
        ctx.save_for_backward(gen, field)
        return K.warp_mix(gen, field)

    @staticmethod
    def backward(ctx, gout):
        gen, field = ctx.saved_tensors
        ggen, gfield = K.warp_mix_bwd(gen, field, gout.contiguous().to(gen.dtype))
        return ggen.to(gen.dtype), gfield.to(field.dtype)


def warp_mix(gen, field):
    """gen NHWC [B,H,W,C] (fp32 / bf16), field fp32 [B,3,H,W] = (dx, dy, alpha) -> NHWC; differentiable in both."""
    return _WarpMix.apply(gen.contiguous(), field)


class _MaskBlend(Function):
    @staticmethod
    def forward(ctx, x, gen, *fields):
        ctx.save_for_backward(x, gen, *fields)
        out, alpha = K.mask_blend(list(fields), x, gen)
        ctx.mark_non_differentiable(alpha)
        return out, alpha

    @staticmethod
    def backward(ctx, gout, _galpha):
        x, gen, *fields = ctx.saved_tensors
        gx, ggen, gfields = K.mask_blend_bwd(fields, x, gen, gout, want_gx=ctx.needs_input_grad[0], want_ggen=ctx.needs_input_grad[1])
        return (gx, ggen, *gfields)


def mask_blend(fields, x, gen):
    """fields: fp32 [B,3,r,r] ascending; x, gen fp32 [B,3,S,S] -> (out, alpha [B,1,S,S]); differentiable in x, gen and the
    alpha channel of every field (alpha itself is returned for the arch's `aligns[1024]` and carries no gradient)."""
    return _MaskBlend.apply(x, gen, *fields)


class _FieldStep(Function):
    @staticmethod
    def forward(ctx, z, prev, coarse, scale):
        ctx.save_for_backward(z, prev, coarse)
        ctx.scale = scale
        return K.field_step(z, prev, coarse, scale)

    @staticmethod
    def backward(ctx, gacc):
        z, prev, coarse = ctx.saved_tensors
        gz, gprev, gcoarse = K.field_step_bwd(z, prev, coarse, gacc, ctx.scale)
        return gz, gprev, gcoarse, None


def field_step(z, prev=None, coarse=None, scale=0.08):
    """z fp32 [B,3,R,R] (AlignNet output before its heads), prev: the field accumulated so far or None, coarse: the coarser
    level's field or None -> accumulated field [B,3,R,R]; differentiable in z, prev and coarse's alpha channel."""
    return _FieldStep.apply(z, prev, coarse, scale)
