"""Per-launch timing of one Adam step of W+ inversion (BASELINE configs[3]: 1024 px, batch 32, bf16) with the convolutions tagged by shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ood_gan_inversion_b200 import kernels as K, stylegan as sg
from ood_gan_inversion_b200.inversion import LatentInverter, generator_synthesizer
from ood_gan_inversion_b200.synth import synthetic_faces, synthetic_init
import ood_gan_inversion_b200.synthesis_grad as SG
sg.set_precision('bf16')
B = int(os.environ.get('B', 32))
dev = 'cuda'
orig = K.conv3x3


def tagged(x, weight, cout, transposed=False, **kw):
    base = kw.get('tag') or 'conv'
    kw['tag'] = f'{base} ci{x.shape[3]} co{cout} {x.shape[1]}px form{int(transposed)}' + ('+ys' if kw.get('want_ys') else '') + ('+f32' if kw.get('out_f32') else '')
    return orig(x, weight, cout, transposed=transposed, **kw)


K.conv3x3 = tagged
for m in (sg, SG):
    if hasattr(m, 'K'):
        m.K.conv3x3 = tagged
torch.manual_seed(0)
gen = synthetic_init(sg.Generator(1024, 512, 8), seed=0).to(dev)
for p in gen.parameters():
    p.requires_grad_(False)
target = synthetic_faces(B, 1024, seed=3, device=dev)
lat0 = torch.zeros(B, 18, 512, device=dev)
inv = LatentInverter(generator_synthesizer(gen), lr=0.01)
inv.run(target, lat0, 2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
inv.run(target, lat0, 3)
e1.record()
torch.cuda.synchronize()
print(f'{e0.elapsed_time(e1) / 3:.2f} ms per Adam step, eager launches (batch {B})')


inv.run(target, lat0, 28)
tm = inv.timing
if tm['replay_steps']:
    print(f"{tm['replay_ms'] / tm['replay_steps']:.2f} ms per Adam step replayed as a CUDA graph ({tm['replay_steps']} replays after {tm['eager_steps']} eager steps)")
K.profile_begin()
inv.run(target, lat0, 1)
torch.cuda.synchronize()
prof = K.profile_end()
tot = sum(v['ms'] for v in prof.values())
print(f'sum of library kernel times {tot:.3f} ms')
for name, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms'])[:45]:
    rate = v['work'] / v['ms'] / 1e9 if v['ms'] > 0 else 0
    print(f"{v['ms']:8.3f} ms {100 * v['ms'] / tot:5.1f}% {v['launches']:4d}x {rate:10.1f} {'TFLOP/s' if 'conv' in name else 'GB/s'}  {name}")
