"""Diagnosis of tests/test_zz_samm_bwd_gpu.py::test_inversion_with_the_blend_in_the_loop (round-1 VERDICT item 1):
first-step dL/dW+ through generator + mask blend vs torch.autograd of the oracle, and where the Adam trajectories part."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
import ood_gan_inversion_b200.stylegan as sg
from ood_gan_inversion_b200.inversion import LatentInverter, blended_synthesizer
from oracle import stylegan as ostyle, samm as osamm
DEV = 'cuda'
sg.set_precision('fp32')
size, batch, steps = 32, 2, 6
rnd = lambda *s, seed: torch.randn(*s, generator=torch.Generator().manual_seed(seed))
sd = ostyle.synthetic_generator_state(size, seed=3)
gen = sg.Generator(size, 512, 8).to(DEV)
gen.load_state_dict(sd)
for p in gen.parameters():
    p.requires_grad_(False)
sdd = {k: v.to(DEV) for k, v in sd.items()}
lat0 = 0.5 * torch.randn(batch, gen.n_latent, 512, generator=torch.Generator().manual_seed(4)).to(DEV)
x = rnd(batch, 3, size, size, seed=6).to(DEV)
target = rnd(batch, 3, size, size, seed=7).to(DEV)
fields = [torch.rand(batch, 3, r, r, generator=torch.Generator().manual_seed(20 + r)).to(DEV) for r in (4, 8)]
ours = blended_synthesizer(gen, fields, x)


def oracle_synth(latent):
    img = ostyle.generator_forward(sdd, latent, size, randomize_noise=False)
    return osamm.blend(osamm.compose_masks(fields, size), x, img)


def grad(f, lat0):
    lat = lat0.clone().requires_grad_(True)
    out = f(lat)
    g, = torch.autograd.grad(F.mse_loss(out, target), lat)
    return g, out.detach()


g, o = grad(ours, lat0)
gr, orr = grad(oracle_synth, lat0)
print('forward max-abs', float((o - orr).abs().max()))
print('grad rel-L2', float((g - gr).norm() / gr.norm()), 'max-abs', float((g - gr).abs().max()), '|g| max', float(gr.abs().max()),
      '|g| median', float(gr.abs().median()))
# plain generator gradient for scale
gp, _ = grad(lambda l: gen(l, input_is_tensor=True, input_is_latent=True, randomize_noise=False)[0], lat0)
gpr, _ = grad(lambda l: ostyle.generator_forward(sdd, l, size, randomize_noise=False), lat0)
print('plain gen grad rel-L2', float((gp - gpr).norm() / gpr.norm()), '|g| max', float(gpr.abs().max()), 'median', float(gpr.abs().median()))
# non-contiguous grad_output into SynthesisFn.backward
lat = lat0.clone().requires_grad_(True)
img = gen(lat, input_is_tensor=True, input_is_latent=True, randomize_noise=False)[0]
go = rnd(batch, size, size, 3, seed=9).to(DEV).permute(0, 3, 1, 2)          # NHWC-strided view
g_nc, = torch.autograd.grad(img, lat, go)
lat2 = lat0.clone().requires_grad_(True)
img2 = gen(lat2, input_is_tensor=True, input_is_latent=True, randomize_noise=False)[0]
g_c, = torch.autograd.grad(img2, lat2, go.contiguous())
print('non-contiguous grad_output: max diff', float((g_nc - g_c).abs().max()), 'of', float(g_c.abs().max()))
# trajectories
lat_a, la = LatentInverter(ours, lr=0.01).run(target, lat0, steps)
lat_b, lb = LatentInverter(oracle_synth, lr=0.01).run(target, lat0, steps)
d = (lat_a - lat_b).abs()
print('losses', la, lb)
print('final latent diff max', float(d.max()), 'mean', float(d.mean()), 'n > 5e-3:', int((d > 5e-3).sum()), 'of', d.numel())
idx = (d > 5e-3).nonzero()
for i in idx[:10]:
    i = tuple(int(v) for v in i)
    print('  elem', i, 'diff', float(d[i]), 'first-step grad ours', float(g[i]), 'oracle', float(gr[i]))
# oracle vs oracle in fp64: how far does Adam amplify fp32 noise by itself?
sdd64 = {k: v.double() for k, v in sdd.items()}
f64 = lambda l: osamm.blend(osamm.compose_masks([f.double() for f in fields], size), x.double(),
                            ostyle.generator_forward(sdd64, l, size, randomize_noise=False))
try:
    lat64 = lat0.double().clone().requires_grad_(True)
    opt = torch.optim.Adam([lat64], lr=0.01)
    for _ in range(steps):
        opt.zero_grad()
        F.mse_loss(f64(lat64), target.double()).backward()
        opt.step()
    d64 = (lat_b.double() - lat64.detach()).abs()
    print('oracle fp32 vs oracle fp64 final latent diff max', float(d64.max()), 'mean', float(d64.mean()))
    d64o = (lat_a.double() - lat64.detach()).abs()
    print('ours fp32 vs oracle fp64 final latent diff max', float(d64o.max()), 'mean', float(d64o.mean()))
except Exception as e:
    print('fp64 oracle failed:', repr(e))
