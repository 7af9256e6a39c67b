"""Per-shape timing of the convolutions of the fast E4E encoder path (batch 16)."""
import os
import sys
import collections

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ood_gan_inversion_b200 import kernels as K, encoder_fast  # noqa: E402
from ood_gan_inversion_b200.encoder import Encoder4Editing  # noqa: E402

torch.manual_seed(0)
enc = Encoder4Editing(50, 'ir_se', {'stylegan_size': 1024}).cuda().eval()
fast = encoder_fast.FastEncoder(enc)
x = torch.randn(16, 3, 256, 256, device='cuda').clamp(-1, 1)
orig = K.conv3x3


def tagged(x, weight, cout, transposed=False, **kw):
    kw['tag'] = f'conv ci{x.shape[3]} co{cout} {x.shape[1]}px s{2 if int(transposed) == 3 else 1}'
    return orig(x, weight, cout, transposed=transposed, **kw)


K.conv3x3 = tagged
encoder_fast.K.conv3x3 = tagged
for _ in range(3):
    fast(x, return_feats=True)
torch.cuda.synchronize()
K.profile_begin()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
fast(x, return_feats=True)
e1.record()
torch.cuda.synchronize()
prof = K.profile_end()
print('encoder forward (with per-launch events):', e0.elapsed_time(e1), 'ms')
tot = 0
for name, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms']):
    tot += v['ms']
    work = v.get('work', 0)
    rate = work / v['ms'] / 1e9 if v['ms'] > 0 else 0
    print(f"{v['ms']:8.3f} ms {v['launches']:4d}x  {rate:9.1f} G/s  {name}")
print('sum', tot)

# the encoder alone as a CUDA graph (how it runs inside the step): ms per replay
K.conv3x3 = orig
encoder_fast.K.conv3x3 = orig
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3):
        fast(x, return_feats=True)
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        out = fast(x, return_feats=True)
    for _ in range(5):
        g.replay()
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(20):
            g.replay()
        e1.record(s)
        s.synchronize()
        ts.append(e0.elapsed_time(e1) / 20)
ts.sort()
print(f'encoder as a graph: best {ts[0]:.3f} ms, median {ts[len(ts) // 2]:.3f} ms per replay (batch 16)')
