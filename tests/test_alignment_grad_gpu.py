"""GPU parity of the differentiable alignment path ("grid_sample grads" and everything between them):
diff_ops Functions against torch.autograd of the same formulas, the AlignNet / SPM_Warp gradient against autograd of the oracle
(oracle/samm.py restates SAMM/helpers.py:62-179), and dL/dW+ of generator + alignment callback + mask blend against autograd
through the oracle pipeline -- same device, TF32 off, fp32 storage (the bf16 route is checked for direction)."""
import types

import pytest
import torch
import torch.nn.functional as F

from oracle import ood as oood, samm as osamm, stylegan as ostyle

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(autouse=True)
def _setup():
    import ood_gan_inversion_b200.stylegan as sg
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sg.set_precision('fp32')
    with torch.enable_grad():
        yield
    sg.set_precision('bf16')


def rnd(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2)


def close(a, b, rtol=1e-4, atol=1e-5, scale=None):
    atol = atol if scale is None else atol * float(scale.abs().max())
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol)


def test_diff_ops_against_autograd():
    from ood_gan_inversion_b200 import diff_ops as D
    b, c, h, w = 2, 32, 9, 11
    x0, g0 = rnd(b, c, h, w, seed=1), rnd(b, c, h, w, seed=2)
    wt, bs = (1 + 0.3 * rnd(c, seed=3)).to(DEV).requires_grad_(True), rnd(c, seed=4).to(DEV).requires_grad_(True)
    # InstanceNorm, affine and not: gx, gw, gb
    for affine in (True, False):
        xr = x0.to(DEV).requires_grad_(True)
        ref = osamm.instance_norm(xr, wt if affine else None, bs if affine else None)
        gr = torch.autograd.grad(ref, [xr] + ([wt, bs] if affine else []), g0.to(DEV))
        xo = nhwc(x0).to(DEV).requires_grad_(True)
        out = D.inst_norm(xo, wt if affine else None, bs if affine else None)
        close(nchw(out), ref.detach(), 1e-4, 1e-5)
        go = torch.autograd.grad(out, [xo] + ([wt, bs] if affine else []), nhwc(g0).to(DEV))
        close(nchw(go[0]), gr[0], 1e-3, 1e-5)
        if affine:
            close(go[1], gr[1], 1e-3, 1e-4)
            close(go[2], gr[2], 1e-3, 1e-4)
    # conv 3x3 / 1x1 data gradient
    for kind, k in (('3x3', 3), ('1x1', 1)):
        wc = (0.1 * rnd(48, c, k, k, seed=5)).to(DEV)
        xr = x0.to(DEV).requires_grad_(True)
        ref = F.conv2d(xr, wc, padding=k // 2)
        gy = rnd(b, 48, h, w, seed=6).to(DEV)
        gr, = torch.autograd.grad(ref, xr, gy)
        xo = nhwc(x0).to(DEV).requires_grad_(True)
        out = D.conv(xo, wc, kind)
        close(nchw(out), ref.detach(), 1e-4, 1e-4)
        go, = torch.autograd.grad(out, xo, nhwc(gy))
        close(nchw(go), gr, 1e-4, 1e-4)
    # PReLU (a negative slope included), add / sub / cat
    slope = torch.tensor([0.25, -0.1] * (c // 2)).to(DEV)
    xr = x0.to(DEV).requires_grad_(True)
    gr, = torch.autograd.grad(F.prelu(xr, slope), xr, g0.to(DEV))
    xo = nhwc(x0).to(DEV).requires_grad_(True)
    out = D.prelu(xo, slope)
    close(nchw(out), F.prelu(x0.to(DEV), slope), 1e-6, 1e-6)
    close(nchw(torch.autograd.grad(out, xo, nhwc(g0).to(DEV))[0]), gr, 1e-6, 1e-6)
    a1, a2 = nhwc(x0).to(DEV).requires_grad_(True), nhwc(g0).to(DEV).requires_grad_(True)
    comb = D.cat(D.sub(a1, a2), D.add(a1, a2))
    assert comb.shape == (b, h, w, 2 * c)
    close(comb, torch.cat([a1 - a2, a1 + a2], -1).detach(), 1e-6, 1e-6)
    gg = nhwc(rnd(b, 2 * c, h, w, seed=7)).to(DEV)
    g1, g2 = torch.autograd.grad(comb, [a1, a2], gg)
    close(g1, gg[..., :c] + gg[..., c:], 1e-6, 1e-6)
    close(g2, -gg[..., :c] + gg[..., c:], 1e-6, 1e-6)
    # the 2C -> 3 head and the 1x1 shortcut
    w27 = (0.1 * rnd(3, c, 3, 3, seed=8)).to(DEV)
    w1 = (0.1 * rnd(3, c, 1, 1, seed=9)).to(DEV).requires_grad_(True)
    g3 = rnd(b, 3, h, w, seed=10).to(DEV)
    xr = x0.to(DEV).requires_grad_(True)
    ref_h, ref_s = F.conv2d(xr, w27, padding=1), F.conv2d(xr, w1)
    gr_h, = torch.autograd.grad(ref_h, xr, g3)
    gr_s = torch.autograd.grad(ref_s, [xr, w1], g3)
    xo = nhwc(x0).to(DEV).requires_grad_(True)
    oh, os_ = D.head27(xo, w27), D.shortcut3(xo, w1)
    close(oh, ref_h.detach(), 1e-4, 1e-4)
    close(os_, ref_s.detach(), 1e-4, 1e-4)
    close(nchw(torch.autograd.grad(oh, xo, g3)[0]), gr_h, 1e-4, 1e-4)
    go_s = torch.autograd.grad(os_, [xo, w1], g3)
    close(nchw(go_s[0]), gr_s[0], 1e-4, 1e-4)
    close(go_s[1], gr_s[1], 1e-3, 1e-3)


def _samm_block(c, scale=0.08, cycles=2, seed=0):
    """A StyledscaleNshfitBlock with xavier-normal convs and perturbed norms + the oracle's state dict of it."""
    from ood_gan_inversion_b200.samm import StyledscaleNshfitBlock
    torch.manual_seed(seed)
    blk = StyledscaleNshfitBlock(c, c, 512, scale=scale, btn=None, cycle_align=cycles, diff_fAndg=True).to(DEV)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in blk.named_parameters():
            if p.dim() == 1 and 'alignment.body' in n and p.numel() > 1:
                if n.endswith('.weight') and 'res_layer.2' not in n:
                    p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
                elif n.endswith('.bias'):
                    p.copy_(0.1 * torch.randn(p.shape, generator=g))
    for p in blk.parameters():
        p.requires_grad_(False)
    sd = {k: v.detach() for k, v in blk.state_dict().items()}
    return blk, sd


@pytest.mark.parametrize('c,r,coarse', [(32, 12, False), (64, 16, True)])
def test_spm_warp_gradient_vs_oracle_autograd(c, r, coarse):
    """SPM_Warp.forward_nhwc_diff (two alignment cycles: AlignNet -> field step -> warp + alpha mix) against torch.autograd of
    oracle.samm.spm_warp: aligned features, field, and the gradients w.r.t. the generator features and the coarser level's field."""
    blk, sd = _samm_block(c)
    b = 2
    gen0, enc0 = rnd(b, c, r, r, seed=1), rnd(b, c, r, r, seed=2)
    coarse0 = torch.rand(b, 3, r // 2, r // 2, generator=torch.Generator().manual_seed(3)) if coarse else None
    g_al, g_f = rnd(b, c, r, r, seed=4).to(DEV), rnd(b, 3, r, r, seed=5).to(DEV)
    gen_r = gen0.to(DEV).requires_grad_(True)
    co_r = coarse0.to(DEV).requires_grad_(True) if coarse else None
    al_r, f_r = osamm.spm_warp(sd, 'alignment.', enc0.to(DEV), gen_r, co_r, 0.08, 2)
    gr = torch.autograd.grad([al_r, f_r], [gen_r] + ([co_r] if coarse else []), [g_al, g_f])
    gen_o = nhwc(gen0).to(DEV).requires_grad_(True)
    co_o = coarse0.to(DEV).requires_grad_(True) if coarse else None
    al_o, f_o = blk.forward_nhwc(enc0.to(DEV), gen_o, co_o)
    assert al_o.requires_grad and f_o.requires_grad
    close(nchw(al_o).detach(), al_r.detach(), 1e-3, 1e-4)
    close(f_o.detach(), f_r.detach(), 1e-3, 1e-5)
    go = torch.autograd.grad([al_o, f_o], [gen_o] + ([co_o] if coarse else []), [nhwc(g_al), g_f])
    rel = float((nchw(go[0]) - gr[0]).norm() / gr[0].norm())
    print(f'SPM_Warp C={c} R={r}: d/dgen rel-L2 {rel:.3g}')
    assert rel < 2e-3
    if coarse:
        close(go[1], gr[1], 1e-2, 1e-3, scale=gr[1])


def _pipeline_case(size=64, batch=1):
    """Generator(size) + two alignment levels (32 and 64 px) + mask blend, ours and the oracle's, as functions of the W+ latents."""
    import ood_gan_inversion_b200.stylegan as sg
    from ood_gan_inversion_b200 import samm_grad
    from ood_gan_inversion_b200.arch import _AlignCallback
    sdg = ostyle.synthetic_generator_state(size, seed=5)
    gen = sg.Generator(size, 512, 8).to(DEV)
    gen.load_state_dict(sdg)
    for p in gen.parameters():
        p.requires_grad_(False)
    sdg = {k: v.to(DEV) for k, v in sdg.items()}
    ch = {32: 512, 64: 512}
    blocks, sds = {}, {}
    for ind, r in ((1, 32), (2, 64)):
        blocks[ind], sds[ind] = _samm_block(ch[r], seed=10 + ind)
    enc = {ind: (0.5 * rnd(batch, ch[r], r, r, seed=20 + ind)).to(DEV) for ind, r in ((1, 32), (2, 64))}
    x = rnd(batch, 3, size, size, seed=30).clamp(-1, 1).to(DEV)
    noises = [rnd(batch, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2), seed=40 + i).to(DEV) for i in range(gen.num_layers)]
    cond_layers = [5, 7]

    owner = types.SimpleNamespace(feats=[enc[2], enc[1]], modulation=[blocks[2], blocks[1]], aligns={}, strict_rng=False)
    cb = _AlignCallback(owner)

    def ours(lat):
        owner.aligns = {}
        conditions = [[None, noises[5]], [None, noises[7]]]        # fixed noise at the conditioned layers too
        # conditions[ci][1] given => model.py:558-571 passes it as that layer's noise and the callback is not consulted; the callback
        # route needs it None, so the seeded draw below supplies the noise
        torch.manual_seed(77)
        img, _ = gen(lat, input_is_tensor=True, input_is_latent=True, noise=noises, conditions=[[None, None], [None, None]],
                     cond_layers=cond_layers, cond_type='NOISE', callback=cb)
        fields = [owner.aligns[1], owner.aligns[2]]
        return samm_grad.mask_blend(fields, x, img)[0], fields

    def oracle(lat):
        aligns = {}

        def hook(ci, image, noise, nw, style):
            ind = ci + 1
            aligned, field = osamm.spm_warp(sds[ind], 'alignment.', enc[ind], image, aligns.get(ind - 1), 0.08, 2)
            aligns[ind] = field
            return (aligned - image + noise * nw) / nw
        torch.manual_seed(77)
        img = ostyle.generator_forward(sdg, lat, size, noise=noises, cond_layers=cond_layers, hook=hook)
        fields = [aligns[1], aligns[2]]
        return osamm.blend(osamm.compose_masks(fields, size), x, img), fields
    lat0 = (0.5 * rnd(batch, gen.n_latent, 512, seed=50)).to(DEV)
    target = rnd(batch, 3, size, size, seed=51).to(DEV)
    return ours, oracle, lat0, target


def test_latent_gradient_through_alignment_and_blend_fp32():
    """dL/dW+ through generator + alignment callback (2 levels x 2 cycles) + mask pyramid + blend == autograd of the oracle."""
    ours, oracle, lat0, target = _pipeline_case()
    la, lb = lat0.clone().requires_grad_(True), lat0.clone().requires_grad_(True)
    oa, fa = ours(la)
    ob, fb = oracle(lb)
    assert oa.requires_grad
    assert float((oa.detach() - ob.detach()).abs().max()) < 1e-3
    for a, b_ in zip(fa, fb):
        assert float((a.detach() - b_.detach()).abs().max()) < 1e-3
    g, = torch.autograd.grad(F.mse_loss(oa, target), la)
    g_r, = torch.autograd.grad(F.mse_loss(ob, target), lb)
    rel = float((g - g_r).norm() / g_r.norm())
    print(f'dL/dW+ through alignment + blend: rel-L2 {rel:.3g}, max-abs {float((g - g_r).abs().max()):.3g} of {float(g_r.abs().max()):.3g}')
    assert rel < 5e-3
    torch.testing.assert_close(g, g_r, rtol=5e-2, atol=5e-3 * float(g_r.abs().max()))


def test_latent_gradient_through_alignment_bf16_direction():
    import ood_gan_inversion_b200.stylegan as sg
    ours, oracle, lat0, target = _pipeline_case()
    lb = lat0.clone().requires_grad_(True)
    g_r, = torch.autograd.grad(F.mse_loss(oracle(lb)[0], target), lb)
    sg.set_precision('bf16')
    la = lat0.clone().requires_grad_(True)
    oa, _ = ours(la)
    g, = torch.autograd.grad(F.mse_loss(oa, target), la)
    cos = float(F.cosine_similarity(g.flatten(), g_r.flatten(), dim=0))
    print(f'bf16 dL/dW+ through alignment + blend: cosine {cos:.4f}')
    assert cos > 0.97


def test_arch_forward_builds_the_graph_like_the_reference():
    """ood_faceGAN_e4e.forward outside no_grad (e4e_arch.py:146-151,258-259: eval_path_length marks the encoder's codes as requiring
    grad): the output carries a graph to `lats`; under no_grad nothing changes."""
    import ood_gan_inversion_b200.stylegan as sg
    from ood_gan_inversion_b200.arch import ood_faceGAN_e4e
    sg.set_precision('bf16')
    sd = oood.synthetic_ood_state(1024, seed=0)
    net = ood_faceGAN_e4e(out_size=1024, style_dim=512, encoder='E4E', enable_modulation=True, warp_scale=0.08, cycle_align=2,
                          blend_with_gen=True, ModSize=64)
    net.load_state_dict(sd, strict=True)
    net = net.to(DEV).eval()
    for p in net.parameters():
        p.requires_grad_(False)
    assert net.eval_path_length is True
    x = F.interpolate(rnd(1, 3, 64, 64, seed=2), (1024, 1024), mode='bicubic', align_corners=False).clamp(-1, 1).to(DEV)
    out, lats = net(x)
    assert out.requires_grad and lats.requires_grad and sorted(net.aligns) == [1, 2, 1024]
    g, = torch.autograd.grad(out.square().mean(), lats)
    assert g.shape == lats.shape and torch.isfinite(g).all() and float(g.abs().max()) > 0
    with torch.no_grad():
        out2, _ = net(x)
    assert not out2.requires_grad


@pytest.mark.parametrize('case', [dict(b=2, h=16, w=16, ci=64, co=128, taps=9), dict(b=1, h=8, w=64, ci=64, co=128, taps=1),
                                  dict(b=2, h=32, w=32, ci=256, co=256, taps=9), dict(b=1, h=16, w=128, ci=128, co=128, taps=9),
                                  dict(b=3, h=4, w=16, ci=192, co=384, taps=9)])
def test_conv_wgrad_tcgen05(case):
    """ood_conv_wgrad (tcgen05, MN-major operands straight from NHWC, K = all pixels in deterministic slices) == autograd's weight
    gradient of conv2d (pad 1 / 1x1) on the same bf16-rounded operands; bit-identical from run to run."""
    from ood_gan_inversion_b200 import kernels as K
    b, h, w, ci, co, taps = (case[k] for k in ('b', 'h', 'w', 'ci', 'co', 'taps'))
    g = rnd(b, co, h, w, seed=1).bfloat16().float().to(DEV)
    x = rnd(b, ci, h, w, seed=2).bfloat16().float().to(DEV)
    k = 3 if taps == 9 else 1
    wt = torch.zeros(co, ci, k, k, device=DEV, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), wt, padding=k // 2).backward(g.double())
    gn, xn = nhwc(g).bfloat16(), nhwc(x).bfloat16()
    out = K.conv_wgrad(gn, xn, taps)
    assert out.shape == (co, ci, k, k)
    close(out, wt.grad.float(), 1e-4, 1e-5, scale=wt.grad)
    assert torch.equal(out, K.conv_wgrad(gn, xn, taps))


def test_alignment_parameter_gradients_vs_oracle_autograd():
    """Training side: gradients of every AlignNet parameter (convolutions through ood_conv_wgrad, InstanceNorm affine
    pairs, PReLU slopes, the 3-channel tail) through two alignment cycles against torch.autograd of the oracle."""
    c, r, b = 64, 16, 2
    blk, sd = _samm_block(c, seed=3)
    params = {n: p for n, p in blk.named_parameters() if n.startswith('alignment.body')}
    for p in params.values():
        p.requires_grad_(True)
    gen0, enc0 = rnd(b, c, r, r, seed=1), rnd(b, c, r, r, seed=2)
    g_al, g_f = rnd(b, c, r, r, seed=4).to(DEV), rnd(b, 3, r, r, seed=5).to(DEV)
    sdr = {k: (v.clone().requires_grad_(True) if k in params else v) for k, v in sd.items()}
    al_r, f_r = osamm.spm_warp(sdr, 'alignment.', enc0.to(DEV), gen0.to(DEV), None, 0.08, 2)
    names = sorted(params)
    gr = torch.autograd.grad((al_r * g_al).sum() + (f_r * g_f).sum(), [sdr[n] for n in names])
    gen_o = nhwc(gen0).to(DEV).requires_grad_(True)
    al_o, f_o = blk.forward_nhwc(enc0.to(DEV), gen_o, None)
    ((nchw(al_o) * g_al).sum() + (f_o * g_f).sum()).backward()
    worst = 0.0
    top = max(float(r.norm()) for r in gr)
    for n, ref in zip(names, gr):
        got = params[n].grad
        assert got is not None and got.shape == ref.shape, n
        # body.0.res_layer.4.bias shifts every channel of out0 by a constant, which the two InstanceNorms behind it remove: its true
        # gradient is zero and both sides hold rounding noise (~1e-7 of the largest gradient), hence the absolute floor
        rel = float((got - ref).norm() / (ref.norm() + 1e-4 * top))
        worst = max(worst, rel)
        assert rel < 5e-3, (n, rel, float(ref.norm()), top)
    print(f'AlignNet parameter gradients ({len(names)} tensors): worst rel-L2 {worst:.3g}')


def test_training_step_of_the_arch():
    """training.generator_step on the drop-in arch (bf16 storage, two alignment levels): the reference's fix list leaves `modulation`
    and `feats_conv` trainable (E4E_Face.yml:123-125); every parameter of the two active AlignNets and of the feats_conv layers they
    read gets a finite gradient, and a few Adam steps on a fixed batch reduce the pixel loss."""
    import ood_gan_inversion_b200.stylegan as sg
    from ood_gan_inversion_b200.arch import ood_faceGAN_e4e
    from ood_gan_inversion_b200.training import apply_fix_list, generator_step, GradAllReduce
    sg.set_precision('bf16')
    sd = oood.synthetic_ood_state(1024, seed=0)
    net = ood_faceGAN_e4e(out_size=1024, style_dim=512, encoder='E4E', enable_modulation=True, warp_scale=0.08, cycle_align=2,
                          blend_with_gen=True, ModSize=64, eval_path_length=False)
    net.load_state_dict(sd, strict=True)
    net = net.to(DEV).eval()
    trainable = apply_fix_list(net)
    names = [n for n, _ in trainable]
    assert all(n.startswith(('modulation', 'feats_conv', 'delta_latent')) for n in names)
    x = F.interpolate(rnd(2, 3, 64, 64, seed=2), (1024, 1024), mode='bicubic', align_corners=False).clamp(-1, 1).to(DEV)
    target = (0.8 * x).detach()
    opt = torch.optim.Adam([p for _, p in trainable], lr=2e-3)
    sync = GradAllReduce([p for _, p in trainable])                  # world size 1: a no-op, exercised for its API
    torch.manual_seed(5)
    losses = [float(generator_step(net, x, target, opt, sync=sync)) for _ in range(4)]
    print('training losses', losses)
    active = [n for n in names if n.startswith(('modulation.3.alignment.body', 'modulation.2.alignment.body', 'feats_conv.3', 'feats_conv.2'))]
    got = dict(trainable)
    assert active and all(got[n].grad is not None and torch.isfinite(got[n].grad).all() for n in active)
    assert any(float(got[n].grad.abs().max()) > 0 for n in active if 'res_layer.1.weight' in n)
    assert losses[-1] < losses[0]
