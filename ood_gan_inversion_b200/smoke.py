"""smoke(): one small generator forward (64 px, batch 2) through the C ABI on cuda:0, checked against the oracle."""
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
    print(f'smoke ok: Generator({size}) max-abs vs oracle fp32 {res["fp32"]:.3g}, bf16 {res["bf16"]:.3g}')
    return res
