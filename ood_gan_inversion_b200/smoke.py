"""smoke(): one small invocation of the hot path through the C ABI on cuda:0, checked against the oracle: a generator forward (64 px,
batch 2; fp32 and bf16), one alignment level (AlignNet on the tcgen05 convolutions, field step, warp + alpha mix) and the mask blend."""
import os
import sys

import torch


def run():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import stylegan as ostyle          # checker only
    from . import stylegan as sg
    if not torch.cuda.is_available():
        raise RuntimeError('smoke() needs a CUDA device')
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = 'cuda:0'
    size = 64
    sd = ostyle.synthetic_generator_state(size, seed=size)
    with torch.no_grad():
        gen = sg.Generator(size, 512, 8).to(dev)
        gen.load_state_dict(sd, strict=True)
        lat = torch.randn(2, gen.n_latent, 512, generator=torch.Generator().manual_seed(1)).to(dev)
        ref = ostyle.generator_forward({k: v.to(dev) for k, v in sd.items()}, lat, size, randomize_noise=False)
        res = {}
        for prec, tol in (('fp32', 1e-3), ('bf16', 2e-2)):
            sg.set_precision(prec)
            img, _ = gen(lat, input_is_tensor=True, input_is_latent=True, randomize_noise=False)
            torch.cuda.synchronize()
            err = float((img - ref).abs().max())
            res[prec] = err
            if not err < tol:
                raise RuntimeError(f'smoke: {prec} generator differs from the oracle by {err} (tolerance {tol})')
        sg.set_precision('bf16')
        # one SAMM level (64 channels, 32 px, two cycles) + the invertibility-mask blend, bf16 storage
        from oracle import samm as osamm
        from . import kernels as K
        from .samm import StyledscaleNshfitBlock
        torch.manual_seed(0)
        blk = StyledscaleNshfitBlock(64, 64, 512, scale=0.08, btn=None, cycle_align=2, diff_fAndg=True).to(dev).eval()
        ssd = {k: v.detach() for k, v in blk.state_dict().items()}
        g = torch.Generator().manual_seed(3)
        feat = torch.randn(2, 64, 32, 32, generator=g).to(dev)
        enc = torch.randn(2, 64, 32, 32, generator=g).to(dev)
        al_ref, f_ref = osamm.spm_warp(ssd, 'alignment.', enc, feat, None, 0.08, 2)
        al, field = blk.forward_nhwc(enc, feat.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16), None)
        x = torch.randn(2, 3, 64, 64, generator=g).clamp(-1, 1).to(dev)
        out, _ = K.mask_blend([field], x, ref)
        out_ref = osamm.blend(osamm.compose_masks([f_ref], 64), x, ref)
        torch.cuda.synchronize()
        res['field'] = float((field - f_ref).abs().max())
        res['aligned'] = float((al.float().permute(0, 3, 1, 2) - al_ref).abs().max())
        res['blend'] = float((out - out_ref).abs().max())
        if not (res['field'] < 2e-2 and res['aligned'] < 0.1 and res['blend'] < 3e-2):
            raise RuntimeError(f'smoke: alignment level / blend differ from the oracle: {res}')
    print(f'smoke ok: Generator({size}) max-abs vs oracle fp32 {res["fp32"]:.3g}, bf16 {res["bf16"]:.3g}; alignment level field {res["field"]:.3g}, '
          f'features {res["aligned"]:.3g}, blend {res["blend"]:.3g}')
    return res
