"""GPU parity of the SAMM backward kernels (ood_warp_mix_bwd, ood_mask_blend_bwd) through the C ABI against torch.autograd of
the oracle on the same device.  Named to run LAST: these kernels were written after the round's GPU budget ended (their
per-item bodies are verified on the CPU by tests/test_samm_bwd_cpu.py), so their first GPU run must not gate the rest."""
import pytest
import torch

from oracle import samm as osamm

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def rnd(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('shape', [(2, 16, 12, 12), (1, 64, 32, 20), (2, 8, 9, 7)])
def test_warp_mix_bwd(dtype, shape):
    from ood_gan_inversion_b200 import samm_grad
    b, c, h, w = shape
    gen0 = rnd(b, c, h, w, seed=1).to(dtype).float()
    field0 = torch.cat([0.3 * rnd(b, 2, h, w, seed=2), torch.rand(b, 1, h, w, generator=torch.Generator().manual_seed(3))], 1)
    gout0 = rnd(b, c, h, w, seed=4).to(dtype).float()
    gen_r, field_r = gen0.to(DEV).requires_grad_(True), field0.to(DEV).requires_grad_(True)
    (osamm.warp_mix(gen_r, field_r) * gout0.to(DEV)).sum().backward()
    gen = gen0.permute(0, 2, 3, 1).contiguous().to(dtype).to(DEV).requires_grad_(True)
    field = field0.to(DEV).requires_grad_(True)
    out = samm_grad.warp_mix(gen, field)
    out.backward(gout0.permute(0, 2, 3, 1).contiguous().to(dtype).to(DEV))
    tol = dict(rtol=1e-4, atol=1e-4) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)     # bf16: the returned ggen is rounded
    torch.testing.assert_close(gen.grad.float().permute(0, 3, 1, 2), gen_r.grad, **tol)
    torch.testing.assert_close(field.grad, field_r.grad, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize('size,levels', [(64, (4, 8, 16, 32)), (128, (8, 16, 32)), (20, (3, 5))])
def test_mask_blend_bwd(size, levels):
    from ood_gan_inversion_b200 import samm_grad
    b = 2
    fields0 = [1.6 * torch.rand(b, 3, r, r, generator=torch.Generator().manual_seed(10 + r)) - 0.3 for r in levels]
    x0, gen0, gout = rnd(b, 3, size, size, seed=1), rnd(b, 3, size, size, seed=2), rnd(b, 3, size, size, seed=3).to(DEV)
    ref_in = [t.to(DEV).requires_grad_(True) for t in (x0, gen0, *fields0)]
    (osamm.blend(osamm.compose_masks(ref_in[2:], size), ref_in[0], ref_in[1]) * gout).sum().backward()
    ours = [t.to(DEV).requires_grad_(True) for t in (x0, gen0, *fields0)]
    out, alpha = samm_grad.mask_blend(ours[2:], ours[0], ours[1])
    assert not alpha.requires_grad
    out.backward(gout)
    for a, r in zip(ours, ref_in):
        torch.testing.assert_close(a.grad, r.grad, rtol=1e-4, atol=1e-4)
