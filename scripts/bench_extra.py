"""Extra measurements for profiles/README.md (not the driver's bench line):
  (a) BASELINE config 4 -- optimisation-based W+ inversion at 1024 px: Adam steps/s through this package's forward+backward
      (bf16 tcgen05) vs torch.autograd through the oracle port on the same GPU (the reference's arithmetic as PyTorch ops);
  (b) the full OOD pipeline forward of the oracle port ON THE GPU (what the unmodified reference effectively executes on
      a GPU: ATen/cuDNN, fp32, default TF32 settings) as a context number for the CPU reference arm.
"""
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ood_gan_inversion_b200 import stylegan as sg  # noqa: E402
from ood_gan_inversion_b200.arch import ood_faceGAN_e4e  # noqa: E402
from ood_gan_inversion_b200.synth import synthetic_faces, synthetic_init  # noqa: E402
from oracle import ood as oood, stylegan as ostyle  # noqa: E402  (baseline arm only)

DEV = 'cuda'


def timed(fn, warm, rep):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rep):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / rep


def inversion(batch, steps, precision):
    sg.set_precision(precision)
    torch.manual_seed(0)
    gen = synthetic_init(sg.Generator(1024, 512, 8), seed=0).to(DEV)
    for p in gen.parameters():
        p.requires_grad_(False)
    target = synthetic_faces(batch, 1024, seed=3, device=DEV)
    lat = torch.zeros(batch, 18, 512, device=DEV, requires_grad=True)
    opt = torch.optim.Adam([lat], lr=0.01)
    losses = []

    def step():
        opt.zero_grad(set_to_none=True)
        img, _ = gen(lat, input_is_tensor=True, input_is_latent=True, randomize_noise=False)
        loss = F.mse_loss(img, target)
        loss.backward()
        opt.step()
        losses.append(loss.detach())
    ms = timed(step, 2, steps)
    mem = torch.cuda.max_memory_allocated() / 2 ** 30
    return dict(what='config4_inversion_ours', precision=precision, batch=batch, ms_per_step=ms, steps_per_s=1e3 / ms,
                image_steps_per_s=batch * 1e3 / ms, loss_first=float(losses[0]), loss_last=float(losses[-1]), peak_mem_gib=mem)


def inversion_oracle(batch, steps):
    sd = {k: v.to(DEV) for k, v in ostyle.synthetic_generator_state(1024, seed=0).items()}
    target = synthetic_faces(batch, 1024, seed=3, device=DEV)
    lat = torch.zeros(batch, 18, 512, device=DEV, requires_grad=True)
    opt = torch.optim.Adam([lat], lr=0.01)

    def step():
        opt.zero_grad(set_to_none=True)
        img = ostyle.generator_forward(sd, lat, 1024, randomize_noise=False)
        F.mse_loss(img, target).backward()
        opt.step()
    torch.cuda.reset_peak_memory_stats()
    ms = timed(step, 1, steps)
    return dict(what='config4_inversion_oracle_autograd_gpu', precision='fp32 (default TF32 settings)', batch=batch, ms_per_step=ms,
                image_steps_per_s=batch * 1e3 / ms, peak_mem_gib=torch.cuda.max_memory_allocated() / 2 ** 30)


def pipeline_oracle_gpu(batch):
    torch.manual_seed(0)
    net = synthetic_init(ood_faceGAN_e4e(out_size=1024, warp_scale=0.08, cycle_align=2, ModSize=256), seed=0).eval()
    sd = {k: v.detach().to(DEV) for k, v in net.state_dict().items()}
    x = synthetic_faces(batch, 1024, device=DEV)
    with torch.no_grad():
        ms = timed(lambda: oood.ood_forward(sd, x, strict_rng=False), 1, 3)
    return dict(what='config2_pipeline_oracle_port_on_gpu', precision='fp32 (default TF32 settings)', batch=batch, ms_per_step=ms,
                images_per_s=batch * 1e3 / ms)


if __name__ == '__main__':
    out = []
    for fn, args in [(inversion, (32, 10, 'bf16')), (inversion, (8, 5, 'bf16')), (inversion_oracle, (2, 3)), (pipeline_oracle_gpu, (4,))]:
        try:
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats()
            r = fn(*args)
        except Exception as e:  # noqa: BLE001
            r = dict(what=fn.__name__, args=str(args), error=repr(e)[:300])
        print(json.dumps(r), flush=True)
        out.append(r)
    json.dump(out, open('gpurun_out/bench_extra.json', 'w'), indent=1)
