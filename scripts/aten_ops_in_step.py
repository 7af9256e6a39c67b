"""Which ATen ops (library launches) does one bf16 forward of the pipeline still issue, and from which line of this package?"""
import os, sys, collections, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.utils._python_dispatch import TorchDispatchMode
torch.set_grad_enabled(False)
import ood_gan_inversion_b200.stylegan as sg
from ood_gan_inversion_b200.arch import ood_faceGAN_e4e
from ood_gan_inversion_b200.synth import synthetic_faces, synthetic_init
sg.set_precision('bf16')
net = synthetic_init(ood_faceGAN_e4e(out_size=1024, style_dim=512, encoder='E4E', enable_modulation=True, warp_scale=0.08, cycle_align=2,
                                     blend_with_gen=True, ModSize=256), seed=0).cuda().eval()
x = synthetic_faces(4, 1024, device='cuda')
net(x); net(x)
PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'ood_gan_inversion_b200')
SKIP = ('aten.empty', 'aten.view', 'aten.permute', 'aten.detach', 'aten.slice', 'aten.select', 'aten.reshape', 'aten._unsafe_view', 'aten.expand',
        'aten.unsqueeze', 'aten.squeeze', 'aten.t.', 'aten.transpose', 'aten.alias', 'aten.as_strided', 'aten.empty_like', 'aten.empty_strided',
        'aten.lift_fresh', 'aten.is_pinned', 'aten._local_scalar')
count = collections.Counter()


class Log(TorchDispatchMode):
    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        name = str(func)
        if not name.startswith(SKIP):
            where = '?'
            for fr in reversed(traceback.extract_stack()):
                if fr.filename.startswith(PKG):
                    where = f'{os.path.basename(fr.filename)}:{fr.lineno}'
                    break
            count[(name, where)] += 1
        return func(*args, **(kwargs or {}))


with Log():
    net(x)
for (name, where), n in sorted(count.items(), key=lambda kv: (kv[0][1], kv[0][0])):
    print(f'{n:4d}  {name:40s} {where}')
print('total', sum(count.values()))
