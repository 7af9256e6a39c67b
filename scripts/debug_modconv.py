import sys, math, torch
sys.path.insert(0, '.')
import torch.nn.functional as F
from ood_gan_inversion_b200 import kernels as K, stylegan as m
from oracle import stylegan as ostyle
torch.set_grad_enabled(False)
G = torch.load('tests/golden/modconv.pt', weights_only=False)
c = G['cases'][0]
m.set_precision('fp32')
mod = m.ModulatedConv2d(c['ci'], c['co'], c['k'], 8, **c['kw']).cuda()
mod.load_state_dict(c['sd'])
wp, wsq, mw, mb = mod.packed()
print('shapes', wp.shape, wsq.shape, mw.shape, mb.shape, mod.cin_p, mod.cout_p)
s, d = mod.coeffs(c['style'].cuda())
s_ref = ostyle.equal_linear(c['style'], c['sd']['modulation.weight'], c['sd']['modulation.bias'])
print('s err', (s.cpu()[:, :4] - s_ref).abs().max().item(), 'pad s', s.cpu()[:, 4:].abs().max().item())
w = c['sd']['weight']
wb = mod.scale * w * s_ref.reshape(2, 1, 4, 1, 1)
d_ref = mod.scale * torch.rsqrt(wb.pow(2).sum([2, 3, 4]) + 1e-8)
print('d err', (d.cpu()[:, :3] - d_ref).abs().max().item(), d.cpu()[0, :5], d_ref[0])
xs = m._to_nhwc(c['x'].cuda(), s, mod.cin_p)
xs_ref = (c['x'] * s_ref[:, :, None, None]).permute(0, 2, 3, 1)
print('xs err', (xs.cpu()[..., :4] - xs_ref).abs().max().item(), xs.shape, xs.cpu()[..., 4:].abs().max().item())
y, _ = K.conv3x3(xs, wp, mod.cout_p, impl=1)
raw_ref = F.conv2d(c['x'] * s_ref[:, :, None, None], w[0], padding=1)
print('raw err', (y.cpu()[..., :3].permute(0, 3, 1, 2) - raw_ref).abs().max().item())
y2, _ = K.conv3x3(xs, wp, mod.cout_p, impl=1, d=d)
print('demod err', (y2.cpu()[..., :3].permute(0, 3, 1, 2) - raw_ref * d_ref[:, :, None, None]).abs().max().item())
out = mod(c['x'].cuda(), c['style'].cuda())
print('module err', (out.cpu() - c['y']).abs().max().item())
