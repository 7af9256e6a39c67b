"""GPU parity of the backward path: kernels vs autograd of the oracle formulas, and the W+ latent
gradient of the whole synthesis vs torch.autograd through the oracle generator (same device, TF32 off)."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import ops as oops, stylegan as ostyle

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(autouse=True)
def _setup():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.enable_grad():
        yield


def K():
    from ood_gan_inversion_b200 import kernels
    return kernels


def rnd(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def nhwc(x, dtype=torch.float32):
    return x.permute(0, 2, 3, 1).contiguous().to(dtype).to(DEV)


def nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous().cpu()


@pytest.mark.parametrize('impl', [0, 1])
@pytest.mark.parametrize('case', [dict(b=2, h=9, w=13, ci=64, co=32), dict(b=1, h=17, w=17, ci=32, co=64), dict(b=3, h=5, w=5, ci=128, co=128),
                                  dict(b=1, h=9, w=271, ci=32, co=32)])
def test_strided_gather_conv(impl, case):
    """transposed=2: out[y,x] = sum in[2y+ky, 2x+kx] W[ky,kx] == F.conv2d(stride=2) == data gradient of conv_transpose2d(stride=2)."""
    b, h, w_, ci, co = case['b'], case['h'], case['w'], case['ci'], case['co']
    dt = torch.bfloat16 if impl == 0 else torch.float32
    x, w = rnd(b, ci, h, w_, seed=1).to(dt).float(), (0.2 * rnd(co, ci, 3, 3, seed=2)).to(dt).float()
    y, _ = K().conv3x3(nhwc(x, dt), K().pack_conv_weight(w.to(DEV), dt, impl == 1), co, transposed=2, impl=impl, out_f32=(impl == 0))
    ref = F.conv2d(x.double(), w.double(), stride=2).float()
    assert y.shape[1:3] == ((h - 1) // 2, (w_ - 1) // 2)
    torch.testing.assert_close(nchw(y), ref, rtol=1e-4, atol=2e-3)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('hw', [(10, 14), (70, 90)])        # tile kernel / row-streaming kernel
def test_blur_adjoint(dtype, hw):
    x = rnd(2, 32, hw[0], hw[1], seed=1).to(dtype).float()
    k = oops.fir_kernel([1, 3, 3, 1], 4.0)
    ref = oops.upfirdn2d(x, torch.flip(k, [0, 1]), pad=(2, 2))         # adjoint of the pad-(1,1) blur
    out, _, _ = K().blur_act(nhwc(x, dtype), list(reversed(K().fir_taps(gain=2.0))), act=False, want_img=True, pad=(2, 2))
    assert out.shape[1:3] == (hw[0] + 1, hw[1] + 1)
    tol = dict(rtol=1e-5, atol=1e-5) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(nchw(out), ref, **tol)
    # <blur(t), g> == <t, blur^T(g)>
    t, g = rnd(1, 32, 9, 9, seed=2), rnd(1, 32, 8, 8, seed=3)
    bt, _, _ = K().blur_act(nhwc(t), K().fir_taps(gain=2.0), act=False, want_img=True)
    bg, _, _ = K().blur_act(nhwc(g), list(reversed(K().fir_taps(gain=2.0))), act=False, want_img=True, pad=(2, 2))
    assert abs(float((bt * nhwc(g)).sum()) - float((nhwc(t) * bg).sum())) < 1e-2


def test_act_bwd_dot_torgb_bwd_vs_autograd():
    b, c, h, w = 2, 32, 9, 11
    acc = rnd(b, c, h, w, seed=1).requires_grad_(True)
    d = (0.5 + rnd(b, c, seed=2).abs()).requires_grad_(True)
    bias, noise, nw = rnd(c, seed=3), rnd(b, 1, h, w, seed=4), torch.tensor([0.3])
    y = oops.fused_leaky_relu(acc * d[:, :, None, None] + nw * noise, bias)
    gy = rnd(b, c, h, w, seed=5)
    g_acc_ref, gd_ref = torch.autograd.grad(y, [acc, d], gy)
    g, gd = K().act_bwd(nhwc(gy), nhwc(y.detach()), d.detach().to(DEV), bias.to(DEV), noise.to(DEV), nw.to(DEV))
    torch.testing.assert_close(nchw(g), g_acc_ref, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(gd.cpu(), gd_ref, rtol=1e-3, atol=1e-3)
    a_, x_ = rnd(b, c, h, w, seed=6), rnd(b, c, h, w, seed=7)
    torch.testing.assert_close(K().dot_reduce(nhwc(a_), nhwc(x_)).cpu(), (a_ * x_).sum((2, 3)), rtol=1e-4, atol=1e-4)
    # ToRGB
    yv = rnd(b, c, h, w, seed=8).requires_grad_(True)
    wrgb = rnd(b, 3, c, seed=9).requires_grad_(True)
    rgb = torch.einsum('bchw,bkc->bkhw', yv, wrgb)
    g_rgb, g_in = rnd(b, 3, h, w, seed=10), rnd(b, c, h, w, seed=11)
    gy_ref, gw_ref = torch.autograd.grad(rgb, [yv, wrgb], g_rgb)
    gy_k, gw_k = K().torgb_bwd(g_rgb.to(DEV), wrgb.detach().to(DEV), nhwc(yv.detach()), nhwc(g_in))
    torch.testing.assert_close(nchw(gy_k), gy_ref + g_in, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(gw_k.cpu(), gw_ref, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize('staged', ['0', '2'])
@pytest.mark.parametrize('dt', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('case', [dict(b=2, c=32, h=19, w=23, rgb=True, gin=True), dict(b=3, c=64, h=16, w=16, rgb=False, gin=True),
                                  dict(b=2, c=512, h=8, w=8, rgb=True, gin=False), dict(b=1, c=128, h=33, w=9, rgb=True, gin=True),
                                  dict(b=40, c=32, h=64, w=96, rgb=True, gin=True), dict(b=36, c=64, h=37, w=53, rgb=False, gin=True)])
def test_act_bwd_fused_equals_the_separate_kernels(case, dt, staged, monkeypatch):
    """ood_act_bwd_fused (one pass: scale of the incoming unscaled gradient + ToRGB data gradient + activation backward + the gd / style /
    ToRGB-weight reductions) against torgb_bwd -> act_bwd -> dot_reduce and against autograd (model.py:277-292, 353-372); both loop forms."""
    monkeypatch.setenv('OOD_ABF_STAGED', staged)          # 0: the register form; 2: the bulk-copy ring form on every shape (partial slots, short chunks, many-slot chunks)
    b, c, h, w = (case[k] for k in ('b', 'c', 'h', 'w'))
    acc = rnd(b, c, h, w, seed=1).requires_grad_(True)
    d = (0.5 + rnd(b, c, seed=2).abs()).requires_grad_(True)
    bias, noise, nw = rnd(c, seed=3), rnd(b, 1, h, w, seed=4), torch.tensor([0.3])
    y = oops.fused_leaky_relu(acc * d[:, :, None, None] + nw * noise, bias)
    yq = y.detach().to(dt).float()                                   # the saved activation in the storage type
    g_in = rnd(b, c, h, w, seed=5).to(dt).float() if case['gin'] else None
    s_up = 0.5 + rnd(b, c, seed=6).abs()
    g_rgb, wrgb = (rnd(b, 3, h, w, seed=7), rnd(b, 3, c, seed=8)) if case['rgb'] else (None, None)
    gy = torch.zeros(b, c, h, w)
    if g_in is not None:
        gy = gy + g_in * s_up[:, :, None, None]
    if g_rgb is not None:
        gy = gy + torch.einsum('bkhw,bkc->bchw', g_rgb, wrgb)
    to = lambda t: None if t is None else t.to(DEV)
    g, gd, dot, gw = K().act_bwd_fused(None if g_in is None else nhwc(g_in).to(dt), to(s_up) if g_in is not None else None,
                                       None if g_rgb is None else (to(g_rgb), to(wrgb)), nhwc(yq).to(dt), d.detach().to(DEV), to(bias), to(noise), to(nw))
    # reference: the same formulas on the stored y (gate on its sign, pre-activation re-derived from it)
    gate = torch.where(yq > 0, torch.tensor(2 ** 0.5), torch.tensor(0.2 * 2 ** 0.5))
    gv = gy * gate
    v = yq / gate - nw * noise - bias[None, :, None, None]
    tol = dict(rtol=1e-4, atol=1e-4) if dt == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(nchw(g), gv * d.detach()[:, :, None, None], **tol)
    red = dict(rtol=1e-3, atol=1e-3 * h * w ** 0.5)
    torch.testing.assert_close(gd.cpu(), (gv * v).sum((2, 3)) / d.detach(), **red)
    torch.testing.assert_close(dot.cpu(), (g_in * yq).sum((2, 3)) if g_in is not None else torch.zeros(b, c), **red)
    if g_rgb is not None:
        torch.testing.assert_close(gw.cpu(), torch.einsum('bkhw,bchw->bkc', g_rgb, yq), **red)
    else:
        assert gw is None
    if dt == torch.float32 and g_in is not None and g_rgb is not None:        # and the chain of separate kernels, fp32: same numbers
        gy_k, gw_k = K().torgb_bwd(to(g_rgb), to(wrgb), nhwc(yq), nhwc(g_in * s_up[:, :, None, None]))
        g_k, gd_k = K().act_bwd(gy_k, nhwc(yq), d.detach().to(DEV), to(bias), to(noise), to(nw))
        torch.testing.assert_close(g, g_k, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(gd, gd_k, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(gw, gw_k, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(dot, K().dot_reduce(nhwc(g_in), nhwc(yq)), rtol=1e-5, atol=1e-4)
    assert torch.equal(g, K().act_bwd_fused(None if g_in is None else nhwc(g_in).to(dt), to(s_up) if g_in is not None else None,
                                            None if g_rgb is None else (to(g_rgb), to(wrgb)), nhwc(yq).to(dt), d.detach().to(DEV), to(bias), to(noise), to(nw))[0])


def _latent_grad_case(size, batch, precision):
    import ood_gan_inversion_b200.stylegan as sg
    sg.set_precision(precision)
    sd = ostyle.synthetic_generator_state(size, seed=size)
    gen = sg.Generator(size, 512, 8).to(DEV)
    gen.load_state_dict(sd, strict=True)
    for p in gen.parameters():
        p.requires_grad_(False)
    lat0 = torch.randn(batch, gen.n_latent, 512, generator=torch.Generator().manual_seed(1)).to(DEV)
    target = torch.randn(batch, 3, size, size, generator=torch.Generator().manual_seed(2)).to(DEV) * 0.3
    lat = lat0.clone().requires_grad_(True)
    img, _ = gen(lat, input_is_tensor=True, input_is_latent=True, randomize_noise=False)
    loss = F.mse_loss(img, target)
    g, = torch.autograd.grad(loss, lat)
    sdd = {k: v.to(DEV) for k, v in sd.items()}
    lat_r = lat0.clone().requires_grad_(True)
    img_r = ostyle.generator_forward(sdd, lat_r, size, randomize_noise=False)
    g_r, = torch.autograd.grad(F.mse_loss(img_r, target), lat_r)
    sg.set_precision('bf16')
    return g, g_r, float((img.detach() - img_r.detach()).abs().max())


@pytest.mark.parametrize('size,batch', [(16, 2), (64, 2), (256, 1)])
def test_latent_gradient_fp32_vs_oracle_autograd(size, batch):
    g, g_r, img_err = _latent_grad_case(size, batch, 'fp32')
    assert img_err < 1e-3
    rel = float((g - g_r).norm() / g_r.norm())
    print(f'fp32 dL/dW+ size {size}: rel-L2 {rel:.3g}, max-abs {float((g - g_r).abs().max()):.3g} (|g| max {float(g_r.abs().max()):.3g})')
    assert rel < 1e-3
    torch.testing.assert_close(g, g_r, rtol=1e-2, atol=1e-3 * float(g_r.abs().max()))


@pytest.mark.parametrize('size,batch', [(64, 2), (256, 1)])
def test_latent_gradient_bf16_direction(size, batch):
    g, g_r, img_err = _latent_grad_case(size, batch, 'bf16')
    cos = float(F.cosine_similarity(g.flatten(), g_r.flatten(), dim=0))
    rel = float((g - g_r).norm() / g_r.norm())
    print(f'bf16 dL/dW+ size {size}: cosine {cos:.5f}, rel-L2 {rel:.3g}')
    assert img_err < 2e-2 and cos > 0.99 and rel < 0.1


def test_adam_inversion_loss_curve_matches_oracle():
    """BASELINE config 4 protocol at small size: N Adam steps on W+ from the same start; loss curves and final latents agree."""
    import ood_gan_inversion_b200.stylegan as sg
    sg.set_precision('fp32')
    size, batch, steps = 32, 2, 8
    sd = ostyle.synthetic_generator_state(size, seed=3)
    gen = sg.Generator(size, 512, 8).to(DEV)
    gen.load_state_dict(sd)
    for p in gen.parameters():
        p.requires_grad_(False)
    sdd = {k: v.to(DEV) for k, v in sd.items()}
    lat0 = 0.5 * torch.randn(batch, gen.n_latent, 512, generator=torch.Generator().manual_seed(4)).to(DEV)
    with torch.no_grad():
        target = ostyle.generator_forward(sdd, torch.randn(batch, gen.n_latent, 512, generator=torch.Generator().manual_seed(5)).to(DEV),
                                          size, randomize_noise=False)
    curves = []
    finals = []
    for which in ('ours', 'oracle'):
        lat = lat0.clone().requires_grad_(True)
        opt = torch.optim.Adam([lat], lr=0.01)
        losses = []
        for _ in range(steps):
            opt.zero_grad()
            if which == 'ours':
                img, _ = gen(lat, input_is_tensor=True, input_is_latent=True, randomize_noise=False)
            else:
                img = ostyle.generator_forward(sdd, lat, size, randomize_noise=False)
            loss = F.mse_loss(img, target)
            loss.backward()
            opt.step()
            losses.append(float(loss))
        curves.append(losses)
        finals.append(lat.detach())
    print('loss curves', curves)
    assert curves[0][-1] < curves[0][0]
    torch.testing.assert_close(torch.tensor(curves[0]), torch.tensor(curves[1]), rtol=1e-3, atol=1e-6)
    assert float((finals[0] - finals[1]).abs().max()) < 5e-3
    sg.set_precision('bf16')


def test_inversion_api_per_image_and_shared_delta():
    """ood_gan_inversion_b200.inversion.invert (BASELINE config 4 in one call): per-image W+ codes reproduce the hand-written Adam
    loop above, and the shared-offset mode (arch delta_latent, OOD_faceGAN_e4e_arch.py:126-129) moves ONE [1,n,512] offset for
    every image and matches torch.autograd through the oracle."""
    import ood_gan_inversion_b200.stylegan as sg
    from ood_gan_inversion_b200.inversion import LatentInverter, generator_synthesizer, invert
    sg.set_precision('fp32')
    size, batch, steps = 32, 2, 6
    sd = ostyle.synthetic_generator_state(size, seed=3)
    gen = sg.Generator(size, 512, 8).to(DEV)
    gen.load_state_dict(sd)
    sdd = {k: v.to(DEV) for k, v in sd.items()}
    lat0 = 0.5 * torch.randn(batch, gen.n_latent, 512, generator=torch.Generator().manual_seed(4)).to(DEV)
    with torch.no_grad():
        target = ostyle.generator_forward(sdd, torch.randn(batch, gen.n_latent, 512, generator=torch.Generator().manual_seed(5)).to(DEV),
                                          size, randomize_noise=False)
    lat, losses = invert(gen, target, lat0, steps=steps, lr=0.01)
    assert lat.shape == lat0.shape and losses[-1] < losses[0]
    oracle = LatentInverter(lambda l: ostyle.generator_forward(sdd, l, size, randomize_noise=False), lr=0.01, graph=False)
    lat_o, losses_o = oracle.run(target, lat0, steps)
    torch.testing.assert_close(torch.tensor(losses), torch.tensor(losses_o), rtol=1e-3, atol=1e-6)
    assert float((lat - lat_o).abs().max()) < 5e-3
    inv = LatentInverter(generator_synthesizer(gen), lr=0.01, shared_delta=True)
    lat_d, losses_d = inv.run(target, lat0, steps)
    oracle_d = LatentInverter(lambda l: ostyle.generator_forward(sdd, l, size, randomize_noise=False), lr=0.01, shared_delta=True, graph=False)
    _, losses_od = oracle_d.run(target, lat0, steps)
    assert inv.delta.shape == (1, gen.n_latent, 512) and losses_d[-1] < losses_d[0]
    torch.testing.assert_close(torch.tensor(losses_d), torch.tensor(losses_od), rtol=1e-3, atol=1e-6)
    # Adam normalises every element's step, so elements whose gradient sits at the rounding noise can walk apart by a few lr:
    # the offsets agree on average and never by more than half of the distance walked (6 steps x lr 0.01)
    diff = (inv.delta - oracle_d.delta).abs()
    assert float(diff.mean()) < 5e-4 and float(diff.max()) < 3e-2
    sg.set_precision('bf16')


@pytest.mark.parametrize('shared', [False, True])
def test_inversion_graph_replay_equals_the_eager_loop(shared):
    """LatentInverter replays the Adam step as a CUDA graph after three eager steps (inversion.py): the same kernels in the same order; the
    loss curve and the final codes follow the all-eager run."""
    import ood_gan_inversion_b200.stylegan as sg
    from ood_gan_inversion_b200.inversion import LatentInverter, generator_synthesizer
    sg.set_precision('bf16')
    size, batch, steps = 64, 3, 11
    gen = sg.Generator(size, 512, 8).to(DEV)
    gen.load_state_dict(ostyle.synthetic_generator_state(size, seed=7))
    for p in gen.parameters():
        p.requires_grad_(False)
    lat0 = 0.5 * torch.randn(batch, gen.n_latent, 512, generator=torch.Generator().manual_seed(8)).to(DEV)
    target = (0.3 * torch.randn(batch, 3, size, size, generator=torch.Generator().manual_seed(9))).to(DEV)
    runs = []
    for graph in (False, True):
        inv = LatentInverter(generator_synthesizer(gen), lr=0.01, shared_delta=shared, graph=graph)
        lat, losses = inv.run(target, lat0, steps)
        runs.append((lat.clone(), losses))
    # same kernels in the same order; the two runs still differ in the optimizer: capturable Adam keeps `step` and the bias corrections on the
    # device in fp32, the eager one computes them in Python doubles -- a last-digit difference per update that Adam's normalisation can amplify on
    # elements whose gradient sits at the rounding noise (see test_inversion_api_per_image_and_shared_delta)
    assert len(runs[1][1]) == steps
    torch.testing.assert_close(torch.tensor(runs[1][1]), torch.tensor(runs[0][1]), rtol=1e-3, atol=1e-6)
    diff = (runs[0][0] - runs[1][0]).abs()
    assert float(diff.mean()) < 5e-4 and float(diff.max()) < 3e-2
    assert runs[1][1][-1] < runs[1][1][0]

