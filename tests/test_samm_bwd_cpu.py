"""Backward of the SAMM gather / blend kernels without a GPU: the per-item bodies of csrc/samm_bwd.cuh (the same source the
__global__ kernels of csrc/samm_bwd.cu wrap) are compiled by g++ into a host emulation (tests/emu/samm_bwd_emu.cpp) and checked
against torch.autograd through the oracle (oracle/samm.py: warp_mix, compose_masks, blend)."""
import ctypes as C
import os
import subprocess

import pytest
import torch

from oracle import ops as oops, samm as osamm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp('emu') / 'samm_bwd_emu.so')
    subprocess.run(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', os.path.join(ROOT, 'tests', 'emu', 'samm_bwd_emu.cpp'), '-o', so],
                   check=True)
    return C.CDLL(so)


def _p(t):
    return C.c_void_p(t.data_ptr())


def rnd(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


@pytest.mark.parametrize('shape,groups', [((2, 8, 6, 7), 4), ((1, 5, 9, 4), 1), ((1, 32, 12, 12), 32)])
def test_warp_mix_bwd_vs_autograd(emu, shape, groups):
    b, c, h, w = shape
    gen = rnd(b, c, h, w, seed=1).requires_grad_(True)
    field = torch.cat([0.3 * rnd(b, 2, h, w, seed=2), torch.rand(b, 1, h, w, generator=torch.Generator().manual_seed(3))], 1)
    field.requires_grad_(True)                                           # large flow: taps fall outside the map
    gout = rnd(b, c, h, w, seed=4)
    (osamm.warp_mix(gen, field) * gout).sum().backward()
    nhwc = lambda t: t.detach().permute(0, 2, 3, 1).contiguous()
    g_n, go_n, f_c = nhwc(gen), nhwc(gout), field.detach().contiguous()
    ggen, gfield = torch.zeros(b, h, w, c), torch.zeros(b, 3, h, w)
    emu.emu_warp_mix_bwd(_p(g_n), _p(f_c), _p(go_n), _p(ggen), _p(gfield), b, h, w, c, groups)
    torch.testing.assert_close(ggen.permute(0, 3, 1, 2), gen.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(gfield, field.grad, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('size,levels', [(32, (4, 8, 16, 32)), (20, (3, 5)), (16, (8,))])
def test_mask_blend_bwd_vs_autograd(emu, size, levels):
    b = 2
    gens = [torch.Generator().manual_seed(10 + r) for r in levels]
    # alphas beyond [0, 1] on purpose: the composition leaves the unit interval and the clip gates the gradient
    fields = [(1.6 * torch.rand(b, 3, r, r, generator=g) - 0.3).requires_grad_(True) for r, g in zip(levels, gens)]
    x, gen = rnd(b, 3, size, size, seed=1).requires_grad_(True), rnd(b, 3, size, size, seed=2).requires_grad_(True)
    gout = rnd(b, 3, size, size, seed=3)
    alpha = osamm.compose_masks(fields, size)
    assert float(((alpha == 0) | (alpha == 1)).float().mean()) > 0
    (osamm.blend(alpha, x, gen) * gout).sum().backward()
    n = len(levels)
    fc = [f.detach().contiguous() for f in fields]
    gf = [torch.zeros_like(f) for f in fc]
    gx, ggen = torch.empty_like(gout), torch.empty_like(gout)
    emu.emu_mask_blend_bwd((C.c_void_p * n)(*[f.data_ptr() for f in fc]), (C.c_void_p * n)(*[g.data_ptr() for g in gf]),
                           (C.c_int * n)(*levels), n, _p(x.detach().contiguous()), _p(gen.detach().contiguous()), _p(gout), _p(gx),
                           _p(ggen), b, size)
    torch.testing.assert_close(gx, x.grad, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ggen, gen.grad, rtol=1e-5, atol=1e-6)
    for g, f in zip(gf, fields):
        assert float(g[:, :2].abs().max()) == 0.0                        # only the alpha channel takes part
        torch.testing.assert_close(g, f.grad, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('r,with_prev,with_coarse', [(12, False, False), (12, True, False), (17, True, True), (9, False, True)])
def test_field_step_bwd_vs_autograd(emu, r, with_prev, with_coarse):
    """heads + FIR + accumulate / clip / PRM + coarse PRM (SAMM/helpers.py:62-77,149-166), as oracle/samm.py:spm_warp composes them."""
    b, scale = 2, 0.08
    k1 = torch.tensor([1., 3., 3., 1.])
    k1 = k1 / k1.sum()
    k = oops.fir_kernel([1, 3, 3, 1])
    z = rnd(b, 3, r, r, seed=1).requires_grad_(True)
    U = lambda *shape, seed: torch.rand(*shape, generator=torch.Generator().manual_seed(seed))
    prev = coarse = None
    if with_prev:      # beyond the valid ranges on purpose: the clips gate the gradient
        prev = torch.cat([scale * (2 * U(b, 2, r, r, seed=2) - 1), 1.4 * U(b, 1, r, r, seed=3) - 0.2], 1).requires_grad_(True)
    if with_coarse:
        coarse = (1.4 * U(b, 3, max(r // 2, 2), max(r // 2, 2), seed=4) - 0.2).requires_grad_(True)
    h = torch.cat([torch.tanh(z[:, 0:1]) * scale, torch.tanh(z[:, 1:2]) * scale, torch.sigmoid(z[:, 2:])], 1)
    acc = oops.upfirdn2d(h, k, pad=(2, 1))
    if prev is not None:
        acc = torch.cat([torch.clip(prev[:, 0:1] + acc[:, 0:1], -scale, scale), torch.clip(prev[:, 1:2] + acc[:, 1:2], -scale, scale),
                         torch.clip(osamm.prm(prev[:, 2:], acc[:, 2:]), 0.0, 1.0)], 1)
    if coarse is not None:
        acc = torch.cat([acc[:, 0:2], torch.clip(osamm.prm(coarse[:, 2:], acc[:, 2:]), 0.0, 1.0)], 1)
    gacc = rnd(b, 3, r, r, seed=5)
    (acc * gacc).sum().backward()
    rc = coarse.shape[-1] if coarse is not None else 0
    gf, gz = torch.zeros(b, 3, r, r), torch.zeros(b, 3, r, r)
    gprev = torch.zeros(b, 3, r, r) if prev is not None else None
    gcoarse = torch.zeros_like(coarse) if coarse is not None else None
    opt = lambda t: _p(t.detach().contiguous()) if t is not None else None
    emu.emu_field_step_bwd(_p(z.detach()), opt(prev), opt(coarse), _p(gacc), _p(k1), C.c_float(scale), b, r, rc, _p(gf), _p(gz),
                           opt(gprev), opt(gcoarse))
    torch.testing.assert_close(gz, z.grad, rtol=1e-4, atol=1e-6)
    if prev is not None:
        torch.testing.assert_close(gprev, prev.grad, rtol=1e-4, atol=1e-6)
    if coarse is not None:
        assert float(gcoarse[:, :2].abs().max()) == 0.0
        torch.testing.assert_close(gcoarse, coarse.grad, rtol=1e-4, atol=1e-5)
