"""Pins the oracle (oracle/) against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import ops, stylegan, samm, ood



@pytest.fixture(autouse=True)
def _no_grad():
    # function-scoped: a module-level torch.set_grad_enabled(False) would leak into every other test module at collection
    with torch.no_grad():
        yield


def test_upfirdn2d_golden(golden):
    for c in golden('ops.pt')['upfirdn2d']:
        y = ops.upfirdn2d(c['x'], c['k'], c['up'], c['down'], c['pad'])
        assert y.shape == c['y'].shape
        torch.testing.assert_close(y, c['y'], rtol=1e-6, atol=1e-6)


def test_upfirdn2d_known_answers():
    # SURVEY.md appendix C (computed from the reference)
    x = torch.arange(16, dtype=torch.float32).view(1, 1, 4, 4)
    k = ops.fir_kernel([1, 3, 3, 1])
    y = ops.upfirdn2d(x, k, pad=(2, 1)).flatten()
    ref = [0.3125, 0.75, 1.25, 1.4375, 1.359375, 2.734375, 3.8125, 3.9375, 3.125, 5.875, 7.5, 7.25,
           4.109375, 7.546875, 9.3125, 8.75]
    torch.testing.assert_close(y, torch.tensor(ref))
    y = ops.upfirdn2d(x, k, down=2, pad=(1, 1)).flatten()
    torch.testing.assert_close(y, torch.tensor([2.734375, 3.9375, 7.546875, 8.75]))
    y = ops.upfirdn2d(x, 4 * k, up=2, pad=(2, 1))
    assert y.shape == (1, 1, 8, 8)
    torch.testing.assert_close(y[0, 0, :3, :3].flatten(),
                               torch.tensor([0, 0.1875, 0.5625, 0.75, 1.25, 1.75, 2.25, 3.25, 3.75]))
    y = ops.upfirdn2d(x, torch.tensor([[.1, .2], [.3, .4]]), pad=(1, 0)).flatten()
    torch.testing.assert_close(y, torch.tensor([0, 0.1, 0.4, 0.7, 0.4, 1.6, 2.6, 3.6, 2.0, 5.6, 6.6, 7.6,
                                                3.6, 9.6, 10.6, 11.6]))


def test_fused_leaky_relu_golden(golden):
    for c in golden('ops.pt')['fused_leaky_relu']:
        torch.testing.assert_close(ops.fused_leaky_relu(c['x'], c['b']), c['y'], rtol=1e-6, atol=1e-7)
    y = ops.fused_leaky_relu(torch.tensor([[-1., 2.], [3., -4.]]), torch.tensor([.5, -.5])).flatten()
    torch.testing.assert_close(y, torch.tensor([-0.14142136, 2.12132025, 4.94974756, -1.27279222]))


def test_fused_leaky_relu_backward_matches_autograd():
    with torch.enable_grad():
        x = torch.randn(2, 5, 4, 4, requires_grad=True)
        b = torch.randn(5, requires_grad=True)
        y = ops.fused_leaky_relu(x, b)
        g = torch.randn_like(y)
        gx, gb = torch.autograd.grad(y, [x, b], g)
    ogx, ogb = ops.fused_leaky_relu_backward(g, y.detach())
    torch.testing.assert_close(ogx, gx)
    torch.testing.assert_close(ogb, gb)


def test_modulated_conv_golden(golden):
    G = golden('modconv.pt')
    for c in G['cases']:
        sd, kw = c['sd'], c['kw']
        up, down = kw.get('upsample', False), kw.get('downsample', False)
        pad = stylegan.up_blur_pads() if up else ((2, 2) if down else None)
        y = stylegan.modulated_conv2d(c['x'], c['style'], sd['weight'], sd['modulation.weight'], sd['modulation.bias'],
                                      kw.get('demodulate', True), up, down, sd.get('blur.kernel'), pad)
        torch.testing.assert_close(y, c['y'], rtol=1e-5, atol=1e-5)
    s = G['styled']
    sd = {k: v for k, v in s['sd'].items()}
    y = stylegan.styled_conv(sd, '', s['x'], s['style'], s['noise'], True)
    torch.testing.assert_close(y, s['y'], rtol=1e-5, atol=1e-5)
    t = G['torgb']
    y = stylegan.to_rgb(t['sd'], '', t['x'], t['style'], t['skip'])
    torch.testing.assert_close(y, t['y'], rtol=1e-5, atol=1e-5)


def test_modulated_conv_kat(golden):
    # torch.manual_seed(0); ModulatedConv2d(4,3,3,8) statistics.  The module's own
    # init draws are replayed in the reference's order (weight, then modulation weight).
    torch.manual_seed(0)
    w = torch.randn(1, 3, 4, 3, 3)
    mw = torch.randn(4, 8)
    x, s = torch.randn(2, 4, 5, 5), torch.randn(2, 8)
    y = stylegan.modulated_conv2d(x, s, w, mw, torch.ones(4))
    kat = golden('modconv.pt')['kat']
    assert abs(float(y.sum()) - kat['sum']) < 1e-4
    assert abs(float(y.abs().sum()) - kat['abssum']) < 1e-4
    assert abs(kat['sum'] - 10.28643799) < 1e-4 and abs(kat['abssum'] - 94.61410522) < 1e-3


@pytest.mark.parametrize('size', [16, 64, 256])
def test_generator_golden(golden, size):
    G = golden('generator.pt')[size]
    sd = stylegan.synthetic_generator_state(size, seed=size)
    lat = torch.randn(G['batch'], G['n_latent'], 512, generator=torch.Generator().manual_seed(1))
    img = stylegan.generator_forward(sd, lat, size, randomize_noise=False)
    st = G['step']
    torch.testing.assert_close(img[:, :, ::st, ::st], G['img'], rtol=1e-4, atol=2e-5)
    assert abs(float(img.double().sum()) - G['img_sum']) < 1e-3 * max(1.0, abs(G['img_abssum'])) * 1e-2
    torch.manual_seed(77)
    img_r = stylegan.generator_forward(sd, lat, size)        # same RNG draw order as the reference
    torch.testing.assert_close(img_r[:, :, ::st, ::st], G['img_rand'], rtol=1e-4, atol=2e-5)
    z = torch.randn(G['batch'], 512, generator=torch.Generator().manual_seed(2))
    torch.testing.assert_close(stylegan.mapping_network(sd, '', z)[:, :16], G['mapping'], rtol=1e-4, atol=1e-5)


def test_samm_golden(golden):
    G = golden('samm.pt')
    sd = {k: v for k, v in G['sd'].items()}
    a1, f1 = samm.spm_warp(sd, 'alignment.', G['enc'], G['gen'], None, 0.08, 2)
    torch.testing.assert_close(f1, G['field'], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(a1, G['aligned'], rtol=1e-4, atol=1e-5)
    a2, f2 = samm.spm_warp(sd, 'alignment.', G['enc2'], G['gen2'], f1, 0.08, 2)
    torch.testing.assert_close(f2, G['field2'], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(a2, G['aligned2'], rtol=1e-4, atol=1e-5)


@pytest.mark.slow
def test_ood_full_pipeline_golden(golden):
    G = golden('ood1024.pt')
    sd = ood.synthetic_ood_state(1024, seed=0)
    x = torch.randn(1, 3, 64, 64, generator=torch.Generator().manual_seed(G['x_small_seed']))
    x = F.interpolate(x, (1024, 1024), mode='bicubic', align_corners=False).clamp(-1, 1)
    torch.manual_seed(123)
    out, lats, aligns = ood.ood_forward(sd, x)
    torch.testing.assert_close(lats, G['lats'], rtol=1e-4, atol=1e-4)
    for k in (1, 2, 3, 4):
        torch.testing.assert_close(aligns[k], G['aligns'][k], rtol=1e-3, atol=2e-4)
    torch.testing.assert_close(aligns[1024][:, :1, ::16, ::16], G['aligns'][1024], rtol=1e-3, atol=2e-4)
    assert (out[:, :, ::16, ::16] - G['out']).abs().max() < 1e-3      # north-star fp32 tolerance
    assert abs(float(out.double().sum()) - G['out_sum']) < 1e-4 * G['out_abssum']


def test_imgio_golden(golden):
    """Byte formats either side of the path (oracle/imgio.py) against the reference's own img2tensor / tensor2img: bit-exact."""
    from oracle import imgio
    G = golden('imgio.pt')
    for f, t in zip(G['frames'], G['frame_tensors']):
        assert torch.equal(imgio.frame_to_tensor(f.numpy()), t)
    for t, f, f01 in zip(G['tensors'], G['tensor_frames'], G['tensor_frames_rgb01']):
        assert torch.equal(torch.from_numpy(imgio.tensor_to_frame(t)), f)
        assert torch.equal(torch.from_numpy(imgio.tensor_to_frame(t, rgb2bgr=False, min_max=(0, 1))), f01)
    # every byte value survives frame -> tensor -> frame
    allv = torch.arange(256, dtype=torch.uint8).repeat_interleave(3).reshape(16, 16, 3).numpy()
    assert (imgio.tensor_to_frame(imgio.frame_to_tensor(allv)) == allv).all()


def test_inversion_golden(golden):
    """BASELINE config 4 protocol (Adam on W+, MSE, frozen weights) through the oracle + torch.autograd against the trajectory of
    the unmodified reference generator: first gradient, loss curve and final latents."""
    G = golden('inversion.pt')
    size, batch, steps = G['size'], G['batch'], G['steps']
    sd = stylegan.synthetic_generator_state(size, seed=3)
    lat0 = 0.5 * torch.randn(batch, G['n_latent'], 512, generator=torch.Generator().manual_seed(4))
    target = stylegan.generator_forward(sd, torch.randn(batch, G['n_latent'], 512, generator=torch.Generator().manual_seed(5)),
                                        size, randomize_noise=False)
    assert float(target.double().sum()) == pytest.approx(G['target_sum'], rel=1e-5, abs=1e-3)
    with torch.enable_grad():
        lat = lat0.clone().requires_grad_(True)
        opt = torch.optim.Adam([lat], lr=0.01)
        losses, grad0 = [], None
        for _ in range(steps):
            opt.zero_grad()
            loss = F.mse_loss(stylegan.generator_forward(sd, lat, size, randomize_noise=False), target)
            loss.backward()
            if grad0 is None:
                grad0 = lat.grad.detach().clone()
            opt.step()
            losses.append(float(loss.detach()))
    rel = float((grad0 - G['grad0']).norm() / G['grad0'].norm())
    assert rel < 1e-4, rel
    torch.testing.assert_close(torch.tensor(losses), torch.tensor(G['losses']), rtol=1e-4, atol=1e-7)
    assert float((lat.detach() - G['final']).abs().max()) < 2e-3
