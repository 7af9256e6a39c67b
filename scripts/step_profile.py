"""Per-launch timing of one bf16 pipeline step (batch 16) with the convolutions tagged by shape: where the step's time goes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_grad_enabled(False)
from ood_gan_inversion_b200 import kernels as K, stylegan as sg
from ood_gan_inversion_b200.arch import ood_faceGAN_e4e
from ood_gan_inversion_b200.synth import synthetic_faces, synthetic_init
sg.set_precision('bf16')
B = int(os.environ.get('B', 16))
net = synthetic_init(ood_faceGAN_e4e(out_size=1024, style_dim=512, encoder='E4E', enable_modulation=True, warp_scale=0.08, cycle_align=2,
                                     blend_with_gen=True, ModSize=256), seed=0).cuda().eval()
x = synthetic_faces(B, 1024, device='cuda')
orig = K.conv3x3.__wrapped__ if hasattr(K.conv3x3, '__wrapped__') else K.conv3x3


def tagged(x, weight, cout, transposed=False, **kw):
    base = kw.get('tag') or 'conv'
    extra = ('+seed' if kw.get('acc_in') is not None else '') + ('+stats' if kw.get('stats_eps') is not None else '') + ('+rgb' if kw.get('rgb') is not None else '') + \
            ('+f32out' if kw.get('out_f32') else '') + (f' g{kw["groups"]}' if kw.get('groups', 1) > 1 else '')
    kw['tag'] = f'{base} ci{x.shape[3]} co{cout} {x.shape[1]}px form{int(transposed)}{extra}'
    return orig(x, weight, cout, transposed=transposed, **kw)


K.conv3x3 = tagged
for _ in range(3):
    net(x)
torch.cuda.synchronize()
K.profile_begin()
net(x)
torch.cuda.synchronize()
prof = K.profile_end()
tot = sum(v['ms'] for v in prof.values())
print(f'sum of kernel times {tot:.3f} ms')
for name, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms']):
    rate = v['work'] / v['ms'] / 1e9 if v['ms'] > 0 else 0
    unit = 'TFLOP/s' if 'conv' in name else 'TB/s'
    print(f"{v['ms']:8.3f} ms {100 * v['ms'] / tot:5.1f}% {v['launches']:4d}x {rate / 1e3 if 'conv' not in name else rate:9.1f} {unit if 'conv' in name else 'GB/s' if False else unit}  {name}")
