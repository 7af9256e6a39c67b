"""Experiment: one CUDA graph that runs the two halves of the batch on two streams, so that HBM-bound kernels of one half overlap the
tensor-bound convolutions of the other.  Prints images/s of the single-stream graph and of the two-stream graph (batch 16)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_grad_enabled(False)
from ood_gan_inversion_b200 import stylegan as sg
from ood_gan_inversion_b200.arch import ood_faceGAN_e4e
from ood_gan_inversion_b200.graphs import GraphedForward
from ood_gan_inversion_b200.synth import synthetic_faces, synthetic_init
sg.set_precision('bf16')
B = int(os.environ.get('B', 16))
net = synthetic_init(ood_faceGAN_e4e(out_size=1024, style_dim=512, encoder='E4E', enable_modulation=True, warp_scale=0.08, cycle_align=2,
                                     blend_with_gen=True, ModSize=256), seed=0).cuda().eval()
x = synthetic_faces(B, 1024, device='cuda')
streams = [torch.cuda.Stream() for _ in range(int(os.environ.get('NS', 2)))]


def split(t):
    cur = torch.cuda.current_stream()
    n = len(streams)
    outs = []
    for i, s in enumerate(streams):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            outs.append(net(t[i * B // n:(i + 1) * B // n])[0])
    for s in streams:
        cur.wait_stream(s)
    return torch.cat(outs)


def bench(fn, name):
    g = GraphedForward(fn, x, warmup=2)
    for _ in range(3):
        g(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f'{name}: {ms:.2f} ms per step, {B / ms * 1e3:.1f} images/s', flush=True)
    return g


bench(lambda t: net(t)[0], 'one stream')
bench(split, f'{len(streams)} streams')
