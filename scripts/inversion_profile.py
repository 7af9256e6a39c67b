"""Per-kernel timing of one Adam step of BASELINE config 4 (W+ inversion, 1024 px, batch 32, bf16)."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ood_gan_inversion_b200 import kernels as K, stylegan as sg  # noqa: E402
from ood_gan_inversion_b200.synth import synthetic_faces, synthetic_init  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sg.set_precision('bf16')
gen = synthetic_init(sg.Generator(1024, 512, 8), seed=0).cuda()
for p in gen.parameters():
    p.requires_grad_(False)
target = synthetic_faces(batch, 1024, seed=3, device='cuda')
lat = torch.zeros(batch, 18, 512, device='cuda', requires_grad=True)
opt = torch.optim.Adam([lat], lr=0.01)


def step():
    opt.zero_grad(set_to_none=True)
    img, _ = gen(lat, input_is_tensor=True, input_is_latent=True, randomize_noise=False)
    loss = F.mse_loss(img, target)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    step()
e1.record()
torch.cuda.synchronize()
print('ms per step (no per-launch events):', e0.elapsed_time(e1) / 5)
K.profile_begin()
step()
torch.cuda.synchronize()
prof = K.profile_end()
tot = 0
for name, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms']):
    tot += v['ms']
    print(f"{v['ms']:8.3f} ms {v['launches']:4d}x  {v['work'] / v['ms'] / 1e9 if v['ms'] else 0:9.1f} G/s  {name}")
print('sum of library kernels', tot)
