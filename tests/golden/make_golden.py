"""Generate golden vectors by importing and running the UNMODIFIED reference (CPU, fp32).

Run in the authoring container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Writes tests/golden/*.pt.  Large weights are never stored: they are regenerated from a seed by
oracle.stylegan.synthetic_generator_state / oracle.ood.synthetic_ood_state and loaded into the
reference modules with strict=True (which also proves key-name coverage).
"""
import os
import sys
import tempfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

shim = tempfile.mkdtemp()
os.makedirs(os.path.join(shim, 'easydict'))
with open(os.path.join(shim, 'easydict', '__init__.py'), 'w') as f:
    f.write("class EasyDict(dict):\n"
            "    def __getattr__(self, k):\n"
            "        try: return self[k]\n"
            "        except KeyError: raise AttributeError(k)\n"
            "    def __setattr__(self, k, v): self[k] = v\n")
sys.path[:0] = ['/root/reference', '/root/reference/BasicSR', shim]

from src.ops.op import upfirdn2d, fused_leaky_relu  # noqa: E402
from src.ops.StyleGAN import model as rm  # noqa: E402
from src.ops.SAMM.helpers import StyledscaleNshfitBlock  # noqa: E402

from oracle.stylegan import synthetic_generator_state  # noqa: E402
from oracle.ood import synthetic_ood_state  # noqa: E402

torch.set_grad_enabled(False)


def save(name, obj):
    path = os.path.join(HERE, name)
    torch.save(obj, path)
    print(f'{name}: {os.path.getsize(path) / 1024:.1f} KiB')


def golden_ops():
    g = torch.Generator().manual_seed(11)
    cases = []
    k4 = rm.make_kernel([1, 3, 3, 1])
    x16 = torch.arange(16, dtype=torch.float32).view(1, 1, 4, 4)
    asym = torch.tensor([[.1, .2], [.3, .4]])
    specs = [
        (x16, k4, 1, 1, (2, 1)), (x16, k4 * 4, 2, 1, (2, 1)), (x16, k4, 1, 2, (1, 1)), (x16, asym, 1, 1, (1, 0)),
        (torch.randn(2, 3, 9, 9, generator=g), k4 * 4, 1, 1, (1, 1)),        # conv-up blur
        (torch.randn(2, 3, 8, 8, generator=g), k4 * 4, 2, 1, (2, 1)),        # rgb skip up2
        (torch.randn(2, 5, 16, 16, generator=g), k4, 1, 2, (2, 2)),          # discriminator down blur
        (torch.randn(1, 2, 7, 5, generator=g), k4, 1, 2, (1, 1)),            # ragged, odd sizes
        (torch.randn(1, 2, 6, 6, generator=g), k4, 1, 1, (-1, 2)),           # negative pad crops
        (torch.randn(1, 1, 5, 5, generator=g), rm.make_kernel([1, 2, 1]), 3, 2, (2, 2)),
        (torch.randn(3, 4, 33, 17, generator=g), k4, 1, 1, (2, 1)),          # field blur, ragged
    ]
    for x, k, up, down, pad in specs:
        cases.append(dict(x=x, k=k, up=up, down=down, pad=pad, y=upfirdn2d(x, k, up=up, down=down, pad=pad)))
    act = []
    for shape in [(2, 2), (2, 6, 5, 7), (3, 4)]:
        x = torch.randn(*shape, generator=g)
        b = torch.randn(shape[1], generator=g)
        act.append(dict(x=x, b=b, y=fused_leaky_relu(x, b)))
    act.append(dict(x=torch.tensor([[-1., 2.], [3., -4.]]), b=torch.tensor([.5, -.5]),
                    y=fused_leaky_relu(torch.tensor([[-1., 2.], [3., -4.]]), torch.tensor([.5, -.5]))))
    save('ops.pt', dict(upfirdn2d=cases, fused_leaky_relu=act))


def golden_modconv():
    out = []
    for seed, (ci, co, k, kw) in enumerate([(4, 3, 3, {}), (6, 8, 3, dict(upsample=True)),
                                            (6, 4, 3, dict(downsample=True)), (8, 3, 1, dict(demodulate=False)),
                                            (16, 16, 3, {}), (16, 8, 3, dict(upsample=True))]):
        torch.manual_seed(seed)
        m = rm.ModulatedConv2d(ci, co, k, 8, **kw)
        x, s = torch.randn(2, ci, 5, 5), torch.randn(2, 8)
        out.append(dict(ci=ci, co=co, k=k, kw=kw, sd=m.state_dict(), x=x, style=s, y=m(x, s)))
    torch.manual_seed(0)
    m = rm.ModulatedConv2d(4, 3, 3, 8)
    y = m(torch.randn(2, 4, 5, 5), torch.randn(2, 8))
    kat = dict(sum=float(y.sum()), abssum=float(y.abs().sum()))
    # StyledConv / ToRGB with non-zero noise weight and biases
    torch.manual_seed(5)
    sc = rm.StyledConv(6, 8, 3, 8, upsample=True)
    sc.noise.weight.fill_(0.3)
    sc.activate.bias.normal_()
    x, s, nz = torch.randn(2, 6, 4, 4), torch.randn(2, 8), torch.randn(2, 1, 8, 8)
    styled = dict(sd=sc.state_dict(), x=x, style=s, noise=nz, y=sc(x, s, noise=nz))
    tr = rm.ToRGB(8, 8)
    tr.bias.normal_()
    x, s, skip = torch.randn(2, 8, 8, 8), torch.randn(2, 8), torch.randn(2, 3, 4, 4)
    rgb = dict(sd=tr.state_dict(), x=x, style=s, skip=skip, y=tr(x, s, skip))
    save('modconv.pt', dict(cases=out, kat=kat, styled=styled, torgb=rgb))


def golden_generator():
    res = {}
    for size, b in [(16, 2), (64, 1), (256, 1)]:
        sd = synthetic_generator_state(size, seed=size)
        gnet = rm.Generator(size, 512, 8)
        gnet.load_state_dict(sd, strict=True)
        gnet.eval()
        n_lat = gnet.n_latent
        lat = torch.randn(b, n_lat, 512, generator=torch.Generator().manual_seed(1))
        img, _ = gnet(lat, input_is_tensor=True, input_is_latent=True, randomize_noise=False)
        torch.manual_seed(77)
        img_rand, _ = gnet(lat, input_is_tensor=True, input_is_latent=True)   # seeded random noise
        z = torch.randn(b, 512, generator=torch.Generator().manual_seed(2))
        wlat = gnet.style(z)
        step = max(1, size // 32)
        res[size] = dict(batch=b, n_latent=n_lat, img=img[:, :, ::step, ::step].clone(), img_sum=float(img.double().sum()),
                         img_abssum=float(img.double().abs().sum()), img_rand=img_rand[:, :, ::step, ::step].clone(),
                         img_min=float(img.min()), img_max=float(img.max()), mapping=wlat[:, :16].clone(), step=step)
        print(size, 'range', float(img.min()), float(img.max()))
    save('generator.pt', res)


def golden_inversion():
    """BASELINE config 4 protocol at small size through the UNMODIFIED reference generator and torch.autograd (which runs the
    reference's own dgrad + grouped wgrad through the materialised per-sample weights): 8 Adam steps on W+ (lr 0.01, MSE), frozen
    weights, registered noise buffers.  Pins the oracle's backward (tests/test_oracle_golden.py::test_inversion_golden), which
    is what the GPU gradient tests compare the hand-written backward with."""
    size, batch, steps = 32, 2, 8
    sd = synthetic_generator_state(size, seed=3)
    gnet = rm.Generator(size, 512, 8)
    gnet.load_state_dict(sd, strict=True)
    gnet.eval()
    for p in gnet.parameters():
        p.requires_grad_(False)
    lat0 = 0.5 * torch.randn(batch, gnet.n_latent, 512, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        target, _ = gnet(torch.randn(batch, gnet.n_latent, 512, generator=torch.Generator().manual_seed(5)),
                         input_is_tensor=True, input_is_latent=True, randomize_noise=False)
    with torch.enable_grad():
        lat = lat0.clone().requires_grad_(True)
        opt = torch.optim.Adam([lat], lr=0.01)
        losses, grad0 = [], None
        for _ in range(steps):
            opt.zero_grad()
            img, _ = gnet(lat, input_is_tensor=True, input_is_latent=True, randomize_noise=False)
            loss = torch.nn.functional.mse_loss(img, target)
            loss.backward()
            if grad0 is None:
                grad0 = lat.grad.detach().clone()
            opt.step()
            losses.append(float(loss.detach()))
    print('inversion losses', losses)
    save('inversion.pt', dict(size=size, batch=batch, steps=steps, n_latent=gnet.n_latent, losses=losses, grad0=grad0,
                              final=lat.detach().clone(), target_sum=float(target.double().sum())))


def golden_samm():
    torch.manual_seed(3)
    blk = StyledscaleNshfitBlock(8, 8, 512, scale=0.08, btn=None, cycle_align=2, diff_fAndg=True)
    for n, p in blk.named_parameters():      # perturb the affine norms away from (1, 0)
        if 'res_layer.0' in n or 'res_layer.4' in n or 'shortcut_layer.1' in n:
            p.add_(0.1 * torch.randn_like(p))
    blk.eval()
    enc, gen = torch.randn(2, 8, 12, 12), torch.randn(2, 8, 12, 12)
    a1, f1 = blk(enc, None, image=gen, aligned=None)
    enc2, gen2 = torch.randn(2, 8, 24, 24), torch.randn(2, 8, 24, 24)
    a2, f2 = blk(enc2, None, image=gen2, aligned=f1)
    save('samm.pt', dict(sd=blk.state_dict(), enc=enc, gen=gen, aligned=a1, field=f1,
                         enc2=enc2, gen2=gen2, aligned2=a2, field2=f2))


def golden_imgio():
    """The byte formats either side of the path: the reference's own functions (basicsr.utils.img_util, loaded from its file:
    the package __init__ pulls in optional dependencies this container lacks) on seeded frames / tensors, incl. every byte
    value, out-of-range values and exact rounding ties."""
    import importlib.util
    import numpy as np
    spec = importlib.util.spec_from_file_location('ref_img_util', '/root/reference/BasicSR/basicsr/utils/img_util.py')
    iu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(iu)
    rng = np.random.default_rng(5)
    frames = [np.arange(256, dtype=np.uint8).repeat(3).reshape(16, 16, 3)[..., ::-1].copy() % np.uint8(251),
              rng.integers(0, 256, (8, 12, 3), dtype=np.uint8), rng.integers(0, 256, (5, 7, 3), dtype=np.uint8)]
    to_t = [((torch.stack(iu.img2tensor([f / 255.0], bgr2rgb=True), dim=0) - 0.5) * 2)[0] for f in frames]   # run_ood_faceGAN_inversion.py:158-159
    g = torch.Generator().manual_seed(6)
    ties = (torch.arange(0, 3 * 8 * 12, dtype=torch.float32).reshape(3, 8, 12) % 256 + 0.5) / 255.0 * 2 - 1  # k + 0.5 before rounding
    tensors = [torch.randn(3, 8, 12, generator=g) * 0.8, torch.randn(3, 5, 7, generator=g) * 2.0, ties]
    # tensor2img clamps a CPU fp32 input IN PLACE (img_util.py:66: .float().detach().cpu() of such a tensor is the tensor itself): pass copies
    to_f = [torch.from_numpy(iu.tensor2img(t.clone(), rgb2bgr=True, min_max=(-1, 1))) for t in tensors]     # :68
    to_f01 = [torch.from_numpy(iu.tensor2img(t.clone(), rgb2bgr=False, min_max=(0, 1))) for t in tensors]
    save('imgio.pt', dict(frames=[torch.from_numpy(f) for f in frames], frame_tensors=to_t, tensors=tensors,
                          tensor_frames=to_f, tensor_frames_rgb01=to_f01))


def golden_ood():
    from src.archs.OOD_faceGAN_e4e_arch import ood_faceGAN_e4e
    sd = synthetic_ood_state(1024, seed=0)
    net = ood_faceGAN_e4e(out_size=1024, style_dim=512, encoder='E4E', enable_modulation=True, warp_scale=0.08,
                          cycle_align=2, blend_with_gen=True, ModSize=256)
    net.load_state_dict(sd, strict=True)
    net.eval()
    gx = torch.Generator().manual_seed(2)
    x = torch.randn(1, 3, 64, 64, generator=gx)
    x = torch.nn.functional.interpolate(x, (1024, 1024), mode='bicubic', align_corners=False).clamp(-1, 1)
    torch.manual_seed(123)
    out, lats = net(x)
    al = net.aligns
    save('ood1024.pt', dict(x_small_seed=2, out=out[:, :, ::16, ::16].clone(), out_sum=float(out.double().sum()),
                            out_abssum=float(out.double().abs().sum()), lats=lats.clone(),
                            aligns={k: (v.clone() if k != 1024 else v[:, :1, ::16, ::16].clone()) for k, v in al.items()},
                            out_min=float(out.min()), out_max=float(out.max())))
    print('ood out range', float(out.min()), float(out.max()), 'alpha mean', float(al[1024].mean()))


if __name__ == '__main__':
    which = sys.argv[1:] or ['ops', 'modconv', 'generator', 'inversion', 'samm', 'imgio', 'ood']
    for w in which:
        globals()[f'golden_{w}']()
