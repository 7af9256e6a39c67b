#!/usr/bin/env python
"""bench.py -- 1024px OOD inversion throughput on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py --gpus 1 --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU arithmetic (oracle port) on host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # one rank per GPU, independent image shards (no collective)

A step = one forward of the full pipeline (E4E encode -> StyleGAN2 1024 synthesis with 4 alignment levels x 2
cycles -> invertibility mask -> ID/OOD blend) over a batch of 16 synthetic 1024x1024 faces per GPU, bf16 storage on
the tcgen05 path (BASELINE.json configs[1]).  `value` times it with the batch resident in HBM; `e2e` times the public
call `net(x)` with pinned-host input, H2D and D2H of the result inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = '1024px inversion images/sec'
UNIT = 'images/s'
BATCH = 16
SIZE = 1024
ARCH_KW = dict(out_size=SIZE, style_dim=512, encoder='E4E', enable_modulation=True, warp_scale=0.08, cycle_align=2,
               blend_with_gen=True, ModSize=256)
WORKLOAD = 'E4E encoder + StyleGAN2 1024px forward + invertibility-mask blend, batch 16 bf16 per GPU (BASELINE configs[1])'


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sus=d.get('bf16_tflops_sustained', d['bf16_tflops']),
                    src='measured (MEASURED_PEAKS.json)')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src='fallback (B200_PROFILING.md)')


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML every 10 ms (nvidia_ml_py), else nvidia-smi."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], False      # rows: (sm_mhz, max_mhz, [reason flags])
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.nvml = pynvml
        except Exception:
            self.nvml = None
        self.t = threading.Thread(target=self.run, daemon=True)

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get('CUDA_VISIBLE_DEVICES', '')
        ids = [v for v in vis.split(',') if v.strip() != '']
        try:
            return int(ids[index]) if ids else index
        except (ValueError, IndexError):
            return index

    def sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(n, 'nvmlDeviceGetCurrentClocksEventReasons') \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        bits = [getattr(n, 'nvmlClocksThrottleReasonHwSlowdown', 0x8), getattr(n, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40),
                getattr(n, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20), getattr(n, 'nvmlClocksThrottleReasonSwPowerCap', 0x4)]
        self.rows.append((float(sm), float(mx), [bool(r & b) for b in bits]))

    def sample_smi(self):
        o = subprocess.run(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits'],
                           capture_output=True, text=True, timeout=5).stdout.strip()
        c = [v.strip() for v in o.split(',')]
        if len(c) >= 7 and c[0].replace('.', '').isdigit():
            self.rows.append((float(c[0]), float(c[1]), [v.lower() == 'active' for v in c[3:7]]))

    def run(self):
        while not self.stop:
            try:
                if self.nvml is not None:
                    self.sample_nvml()
                else:
                    self.sample_smi()
            except Exception:
                if self.nvml is not None:
                    self.nvml = None                     # fall back to nvidia-smi
            time.sleep(0.01 if self.nvml is not None else 0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['unavailable'])
        sm = sorted(r[0] for r in self.rows)
        reasons = [n for i, n in enumerate(self.NAMES) if any(r[2][i] for r in self.rows)]
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=self.rows[0][1], samples=len(sm), reasons=reasons,
                    source='nvml' if self.nvml is not None else 'nvidia-smi')


REF_STAGE = os.path.join(ROOT, 'baseline', '_ref')


def reference_available():
    return os.path.isfile(os.path.join(REF_STAGE, 'src', 'archs', 'OOD_faceGAN_e4e_arch.py'))


def cpu_reference(steps, warmup, images_per_step=1, state=None, prefer_reference=True):
    """The reference's own CPU implementation of the path on the host cores, one bounded sample (`images_per_step` images of
    the same 1024px pipeline) per step.

    kind "reference": the UNMODIFIED reference modules staged under baseline/_ref by __graft_entry__.build() -- its
    `ood_faceGAN_e4e(...)(x)` through its own API with the default device='cpu' branch of its ops (upfirdn2d_native, native
    fused_leaky_relu).  Needs a process that sees no CUDA device (the reference JIT-builds its legacy extensions at import time
    when one is visible, src/ops/op/upfirdn2d.py:10-18): --impl reference hides the GPUs before torch is imported.
    kind "port": oracle/ (plain-PyTorch restatement of the same arithmetic), when the staged copy is absent."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    torch.set_grad_enabled(False)
    from ood_gan_inversion_b200.synth import synthetic_faces, synthetic_init
    x = synthetic_faces(images_per_step, SIZE)
    kind = 'port'
    if prefer_reference and reference_available() and not torch.cuda.is_available():
        for pth in (REF_STAGE,):
            if pth not in sys.path:
                sys.path.insert(0, pth)
        from src.archs.OOD_faceGAN_e4e_arch import ood_faceGAN_e4e as ref_arch
        torch.manual_seed(0)
        net = synthetic_init(ref_arch(**ARCH_KW), seed=0).eval()
        if state is not None:
            net.load_state_dict(state, strict=True)
        run = lambda: net(x)
        kind = 'reference'
    else:
        from oracle import ood as oood
        sd = state if state is not None else oood.synthetic_ood_state(SIZE, seed=0)
        run = lambda: oood.ood_forward(sd, x, size=SIZE, strict_rng=False)
    times = []
    for i in range(warmup + steps):
        torch.manual_seed(123)
        t0 = time.perf_counter()
        run()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    what = 'unmodified reference modules (baseline/_ref), ood_faceGAN_e4e.forward' if kind == 'reference' else 'oracle port'
    return dict(value=images_per_step * len(times) / total, ms_per_step=1e3 * total / len(times), cores=torch.get_num_threads(), kind=kind,
                sample=f'{len(times)} step(s) x {images_per_step} image(s) of the 1024px pipeline ({what}), fp32, after {warmup} warm-up')


def cpu_op_baselines(device=None):
    """SURVEY section 8(d) "CPU baseline timing": the reference's native branch (`upfirdn2d_native`, native
    `fused_leaky_relu`: oracle/ops.py restates src/ops/op/upfirdn2d.py:160-193 and fused_act.py:92-96) and BASELINE
    configs[0] (Generator(256) forward, batch 1, fixed noise) on the host cores -- median of 3 calls after one warm-up,
    one image per call (bounded sample: ~10 s of CPU work).  With `device`, the same op through this library on the
    same shape at batch 4 (CUDA events, best of 10; every tensor is larger than L2 or the case is marked latency).
    Reported baselines only."""
    import torch
    from oracle import ops as oops, stylegan as ostyle
    torch.set_num_threads(os.cpu_count() or 1)
    torch.set_grad_enabled(False)

    def cpu_ms(fn):
        fn()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        return 1e3 * sorted(ts)[1]

    def gpu_ms(fn):
        for _ in range(3):
            fn()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(device)
            best = min(best, e0.elapsed_time(e1))
        return best

    k = oops.fir_kernel([1, 3, 3, 1])
    cases = [('upfirdn2d blur up1 pad(1,1)', (32, 1025, 1025), dict(up=1, down=1, pad=(1, 1)), 4.0),
             ('upfirdn2d up2 pad(2,1)', (3, 512, 512), dict(up=2, down=1, pad=(2, 1)), 4.0),
             ('upfirdn2d down2 pad(1,1)', (32, 1024, 1024), dict(up=1, down=2, pad=(1, 1)), 1.0)]
    rows = []
    for name, chw, kw, gain in cases:
        x = torch.randn(1, *chw)
        y = oops.upfirdn2d(x, k * gain, **kw)
        byt = (x.numel() + y.numel()) * 4                                   # section 8(d): in + out bytes per image
        ms = cpu_ms(lambda: oops.upfirdn2d(x, k * gain, **kw))
        row = dict(op=name, shape=[1, *chw], cpu_ms=ms, cpu_gbs=byt / ms / 1e6)
        if device is not None:
            from ood_gan_inversion_b200 import op as pop
            xg, kg = torch.randn(4, *chw, device=device), (k * gain).to(device)
            g = gpu_ms(lambda: pop.upfirdn2d(xg, kg, **kw))
            row.update(gpu_shape=[4, *chw], gpu_us=1e3 * g, gpu_gbs=4 * byt / g / 1e6)
            del xg
        rows.append(row)
    x, bias = torch.randn(1, 32, 1024, 1024), torch.randn(32)
    ms = cpu_ms(lambda: oops.fused_leaky_relu(x, bias))
    row = dict(op='fused_leaky_relu', shape=[1, 32, 1024, 1024], cpu_ms=ms, cpu_gbs=2 * x.numel() * 4 / ms / 1e6)
    if device is not None:
        from ood_gan_inversion_b200 import op as pop
        xg, bg = torch.randn(4, 32, 1024, 1024, device=device), bias.to(device)
        g = gpu_ms(lambda: pop.fused_leaky_relu(xg, bg))
        row.update(gpu_shape=[4, 32, 1024, 1024], gpu_us=1e3 * g, gpu_gbs=4 * 2 * x.numel() * 4 / g / 1e6)
        del xg
    rows.append(row)
    # BASELINE configs[0]: StyleGAN2 256px synthesis forward, batch 1, W+ latents (seed 1), registered noise buffers
    sd = ostyle.synthetic_generator_state(256, seed=0)
    lat = torch.randn(1, 14, 512, generator=torch.Generator().manual_seed(1))
    ms = cpu_ms(lambda: ostyle.generator_forward(sd, lat, 256, randomize_noise=False))
    row = dict(op='Generator(256) forward, batch 1 (BASELINE configs[0])', shape=[1, 14, 512], cpu_ms=ms, cpu_images_per_s=1e3 / ms)
    if device is not None:
        from ood_gan_inversion_b200 import stylegan as sgm
        gen = sgm.Generator(256, 512, 8).to(device)
        gen.load_state_dict(sd, strict=True)
        gen.eval()
        lg = lat.to(device)
        g = gpu_ms(lambda: gen(lg, input_is_tensor=True, input_is_latent=True, randomize_noise=False))
        row.update(gpu_us=1e3 * g, gpu_images_per_s=1e3 / g, gpu_note='eager launches, batch 1: launch-latency bound')
    rows.append(row)
    return dict(cores=torch.get_num_threads(), kind='port', ops=rows)


def sharded_batch_leg(net, global_batch, chunk, steps, rank, world, dev, barrier, reduce_max, want_e2e=True):
    """BASELINE configs[2] (SURVEY.md section 8d "Config 3"): `global_batch` images split evenly over the ranks
    (sharding.shard_bounds: contiguous slices, no data-path collective), each rank running its slice in micro-batches of
    <= `chunk` images, per-rank seed = base + rank.  images/s = global_batch x passes / max-rank time (CUDA events).
    Returns dict(value=device-resident, e2e=pinned host in/out with the copies inside the timed region)."""
    import torch
    from ood_gan_inversion_b200.graphs import GraphedForward, PipelinedForward
    from ood_gan_inversion_b200.sharding import shard_bounds
    from ood_gan_inversion_b200.synth import synthetic_faces
    lo, hi = shard_bounds(global_batch, world, rank)
    n = hi - lo
    nch = max(1, -(-n // chunk))
    sizes = [n // nch + (1 if i < n % nch else 0) for i in range(nch)] if n > 0 else []
    bounds, a = [], 0
    for sz in sizes:
        bounds.append((a, a + sz))
        a += sz
    x_host = synthetic_faces(max(n, 1), SIZE, seed=100 + rank, pin=True)[:n]
    x_dev = x_host.to(dev)
    out_dev = torch.empty_like(x_dev)
    fn = lambda t: net(t)[0]
    graphs = {}
    for sz in sorted(set(sizes)):
        torch.manual_seed(3000 + rank)
        graphs[sz] = GraphedForward(fn, x_dev[:sz], warmup=2)

    def one_pass():
        for a, b in bounds:
            out_dev[a:b].copy_(graphs[b - a](x_dev[a:b]), non_blocking=True)
    for _ in range(3):
        one_pass()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one_pass()
    e1.record()
    barrier()
    ms = reduce_max(e0.elapsed_time(e1))
    res = dict(global_batch=global_batch, images_per_rank=n, chunk=max(sizes) if sizes else 0, chunks_per_rank=len(sizes), passes=steps,
               value=global_batch * steps / (ms * 1e-3), unit=UNIT, ms_per_pass=ms / steps, scaling='strong',
               note='BASELINE configs[2]: images split by sharding.shard_bounds, no collective; value = global_batch x passes / max-rank time')
    if want_e2e and len(set(sizes)) == 1:
        out_host = torch.empty(n, 3, SIZE, SIZE, dtype=torch.float32).pin_memory()
        torch.manual_seed(3000 + rank)
        pipe = PipelinedForward(fn, x_dev[:sizes[0]], depth=2, warmup=1)

        def host_pass():                      # consecutive batches stream through the two-graph pipeline: no drain between passes
            for a, b in bounds:
                pipe.submit(x_host[a:b], out_host[a:b])
        host_pass()
        pipe.synchronize()
        barrier()
        e0.record()
        for _ in range(steps):
            host_pass()
        pipe.synchronize()
        e1.record()
        barrier()
        ms2 = reduce_max(e0.elapsed_time(e1))
        res['e2e'] = dict(value=global_batch * steps / (ms2 * 1e-3), unit=UNIT, ms_per_pass=ms2 / steps,
                          h2d_bytes_per_pass_per_rank=x_host.numel() * 4, d2h_bytes_per_pass_per_rank=out_host.numel() * 4)
        del pipe, out_host
    del graphs, x_dev, out_dev
    torch.cuda.empty_cache()
    return res


def inversion_leg(dev, batch=32, steps=5):
    """BASELINE configs[3] ("Config 4"): Adam on W+ at 1024 px through this package's forward + hand-written backward
    (inversion.LatentInverter, bf16 storage), `steps` timed steps after a 3-step warm-up run."""
    import torch
    from ood_gan_inversion_b200 import stylegan as sg
    from ood_gan_inversion_b200.inversion import LatentInverter, generator_synthesizer
    from ood_gan_inversion_b200.synth import synthetic_faces, synthetic_init
    with torch.enable_grad():
        torch.manual_seed(0)
        gen = synthetic_init(sg.Generator(SIZE, 512, 8), seed=0).to(dev)
        for p in gen.parameters():
            p.requires_grad_(False)
        target = synthetic_faces(batch, SIZE, seed=3, device=dev)
        lat0 = torch.zeros(batch, 18, 512, device=dev)
        inv = LatentInverter(generator_synthesizer(gen), lr=0.01)
        inv.run(target, lat0, 3)                         # warm-up: kernels loaded, the caching allocator at its steady state
        torch.cuda.synchronize(dev)
        torch.cuda.reset_peak_memory_stats(dev)
        # One run of LatentInverter, timed as a whole with CUDA events (eager launches, the default).  With OOD_INVERSION_GRAPH=1 the run replays the
        # Adam step as a CUDA graph after three eager steps; ms per step is then the device time of the replayed steps (events inside run()).
        n_run = 3 + 4 * steps
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _, losses = inv.run(target, lat0, n_run)
        e1.record()
        torch.cuda.synchronize(dev)
        run_ms = e0.elapsed_time(e1)
        tm = inv.timing
    n_timed = tm['replay_steps'] if tm['replay_steps'] else n_run
    ms = (tm['replay_ms'] if tm['replay_steps'] else run_ms) / n_timed
    res = dict(workload=f'W+ latent inversion, 1024 px, Adam lr 0.01, pixel MSE, batch {batch}, bf16 (BASELINE configs[3])', batch=batch,
               steps=n_timed, ms_per_step=ms, steps_per_s=1e3 / ms, image_steps_per_s=batch * 1e3 / ms, loss_first=losses[0],
               loss_last=losses[-1], peak_mem_gib=torch.cuda.max_memory_allocated(dev) / 2 ** 30,
               timing=('CUDA events around the graph-replayed steps of one run' if tm['replay_steps'] else 'CUDA events around an all-eager run'),
               run=dict(steps=n_run, ms=run_ms, eager_steps=tm['eager_steps'], note='whole run incl. the eager steps and the graph capture'))
    del gen, inv, target
    torch.cuda.empty_cache()
    return res


def gpu_reference_leg(state, dev, batch=4, steps=3):
    """The reference's arithmetic for the same pipeline ON THE SAME B200: the oracle port (plain PyTorch ops: ATen / cuDNN grouped
    convolutions, upfirdn2d_native, unfused elementwise passes -- what the reference's default branch executes, SURVEY finding 2),
    fp32, TF32 off, batch 4.  The like-for-like bar for this library's kernels; the oracle is only the thing timed here."""
    import torch
    from oracle import ood as oood
    from ood_gan_inversion_b200.synth import synthetic_faces
    tf = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        x = synthetic_faces(batch, SIZE, seed=2, device=dev)
        for _ in range(2):
            oood.ood_forward(state, x, size=SIZE, strict_rng=False)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            oood.ood_forward(state, x, size=SIZE, strict_rng=False)
        e1.record()
        torch.cuda.synchronize(dev)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf
    ms = e0.elapsed_time(e1) / steps
    return dict(value=batch * 1e3 / ms, unit=UNIT, ms_per_step=ms, batch=batch, steps=steps, dtype='f32', tf32=False,
                kind='oracle port on the same GPU (ATen/cuDNN), eager launches')


def parity_leg(net, state, x, dev, seed=123):
    """The benched batch against the fp32 oracle on the same device (TF32 off, same seed and RNG call order => identical noise),
    outside every timed region: north_star's bf16 tolerances are max-abs < 2e-2 and PSNR >= 40 dB on the image."""
    import math
    import torch
    from oracle import ood as oood
    tf = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    strict = net.strict_rng
    try:
        net.strict_rng = True
        torch.manual_seed(seed)
        out, lats = net(x)
        aligns = {k: v for k, v in net.aligns.items()}
        torch.manual_seed(seed)
        ref, rlats, raligns = oood.ood_forward(state, x, size=SIZE, strict_rng=True)
        mse = float(((out - ref) ** 2).mean())
        res = dict(batch=int(x.shape[0]), max_abs=float((out - ref).abs().max()), psnr=10 * math.log10(4.0 / max(mse, 1e-30)),
                   lats_max_abs=float((lats - rlats).abs().max()), mask_max_abs=float((aligns[1024] - raligns[1024]).abs().max()),
                   fields_max_abs={str(k): float((aligns[k] - raligns[k]).abs().max()) for k in (1, 2, 3, 4)},
                   tolerance=dict(max_abs=2e-2, psnr=40.0), against='oracle port, fp32, TF32 off, same device, same seed (strict RNG order)')
        res['ok'] = bool(res['max_abs'] < 2e-2 and res['psnr'] >= 40.0)
    finally:
        net.strict_rng = strict
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf
    return res


def fp32_mode_leg(net, state, x, dev, batch=2, steps=2):
    """The fp32 parity mode (SIMT FFMA convolutions, fp32 storage; `set_precision('fp32')`) as a product mode: images/s at a small batch
    and its max-abs against the fp32 oracle (tolerance 1e-3).  Outside every timed region of the headline; eager launches."""
    import torch
    import ood_gan_inversion_b200.stylegan as sg
    from oracle import ood as oood
    tf = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    strict = net.strict_rng
    xb = x[:batch].contiguous()
    try:
        sg.set_precision('fp32')
        net.strict_rng = True
        torch.manual_seed(7)
        out, _ = net(xb)
        torch.manual_seed(7)
        ref, _, _ = oood.ood_forward(state, xb, size=SIZE, strict_rng=True)
        err = float((out - ref).abs().max())
        net.strict_rng = strict
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            net(xb)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return dict(value=batch / (ms * 1e-3), unit=UNIT, batch=batch, steps=steps, ms_per_step=ms, max_abs_vs_oracle=err, tolerance=1e-3,
                    ok=bool(err < 1e-3), note='parity mode: fp32 storage, SIMT FFMA convolutions (csrc/conv_simt.cu), PyTorch encoder modules; eager')
    finally:
        net.strict_rng = strict
        sg.set_precision('bf16')
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf


def main_config(world, batch, launch):
    """`config` of the JSON line (both arms print the same one): BASELINE configs[1]."""
    return dict(workload=WORKLOAD, batch_per_gpu=batch, global_batch=batch * world, size=SIZE, cycle_align=2, mod_size=256,
                l2='inputs (201 MB/step) exceed the 126 MB L2', launch=launch,
                parallelism=f'independent image shards x{world}, no collective',
                weights='random-init (reference init + non-zero noise weights), synthetic smooth faces')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=BATCH)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--ncu', action='store_true', help='run warm-up, then ONE step between cudaProfilerStart/Stop and exit '
                    '(for `ncu --profile-from-start off`; never a bench value)')
    ap.add_argument('--u8-io', action='store_true', help='(default on; kept for compatibility) time the byte-format serving loop')
    ap.add_argument('--no-u8-io', action='store_true', help='skip the byte-format serving loop (imgio.ByteServing: uint8 BGR frames '
                    'across PCIe both ways, 3 instead of 12 bytes per pixel), reported as "e2e_u8"')
    ap.add_argument('--global-batch', type=int, default=0, help='BASELINE configs[2]: this many images per step split over the ranks '
                    '(sharding.shard_bounds) and run in chunks of <= --chunk; value = images / max-rank time (strong scaling). '
                    'Without it the main line is configs[1] (16 images per GPU, weak scaling) and a short configs[2] leg is '
                    'reported under "config3"')
    ap.add_argument('--chunk', type=int, default=32, help='largest per-GPU micro-batch of the sharded batch (SURVEY 8d: <= 32)')
    ap.add_argument('--no-extra-legs', action='store_true', help='skip the config3 / config4 / gpu_reference / parity legs')
    ap.add_argument('--no-graph', action='store_true', help='time eager launches instead of a CUDA-graph replay of the step')
    ap.add_argument('--profile', action='store_true', help='print a torch.profiler kernel table for one step (not a bench value)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))

    if args.impl == 'reference':
        if rank != 0:
            return 0
        os.environ['CUDA_VISIBLE_DEVICES'] = ''                  # the reference's CPU branch; also keeps its import from JIT-building
        r = cpu_reference(max(1, args.steps), max(0, args.warmup))
        line = dict(metric=METRIC, value=r['value'], unit=UNIT, impl='reference', n_gpus=args.gpus, steps=args.steps,
                    warmup=args.warmup, ms_per_step=r['ms_per_step'], higher_is_better=True, scaling='weak',
                    vs_baseline=None, dtype='f32', data='synthetic',
                    config=main_config(args.gpus, BATCH, 'host CPU, all cores'),
                    cpu_baseline=dict(value=r['value'], unit=UNIT, cores=r['cores'], kind=r['kind'], sample=r['sample'],
                                      note='each step is a bounded sample of the workload: 1 image of the same 1024px pipeline '
                                           '(a 16-image step takes ~25 s on these cores); images/s is per image either way'),
                    e2e=dict(value=r['value'], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
        if not args.no_cpu_baseline:
            try:
                line['cpu_baseline']['ops'] = cpu_op_baselines()['ops']
            except Exception as exc:                                 # a reported baseline must not cost the line
                line['cpu_baseline']['ops_error'] = f'{type(exc).__name__}: {exc}'[:300]
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    # stdout carries exactly one JSON line: libraries that print there (NCCL's version banner) are sent to stderr, and the
    # line is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    torch.set_grad_enabled(False)

    from ood_gan_inversion_b200 import _lib, kernels as K, stylegan as sg
    from ood_gan_inversion_b200.arch import ood_faceGAN_e4e
    from ood_gan_inversion_b200.synth import synthetic_faces, synthetic_init
    _lib.lib()                                    # fail loudly before anything else if the CUDA library is missing
    sg.set_precision('bf16')
    torch.manual_seed(0)
    net = synthetic_init(ood_faceGAN_e4e(**ARCH_KW), seed=0).to(dev).eval()
    B = args.batch
    x_host = synthetic_faces(B, SIZE, seed=2 + rank, pin=True)
    x_dev = x_host.to(dev)
    out_host = torch.empty(B, 3, SIZE, SIZE, dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(ms_local):
        t = torch.tensor([ms_local], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.global_batch > 0:
        # BASELINE configs[2] as the main line: strong scaling of one sharded batch (no per-kernel legs, no CPU baseline)
        for _ in range(max(args.warmup, 3)):
            torch.manual_seed(1000 + rank)
            net(x_dev)
        with ClockSampler(local_rank) as clk:
            r = sharded_batch_leg(net, args.global_batch, args.chunk, args.steps, rank, world, dev, barrier, reduce_max)
        if rank == 0:
            line = dict(metric=METRIC, value=r['value'], unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                        ms_per_step=r['ms_per_pass'], higher_is_better=True, scaling='strong', vs_baseline=None, dtype='bf16',
                        data='synthetic', clocks=clk.summary(),
                        config=dict(workload=f'Full OOD inversion inference 1024px, batch {args.global_batch} sharded over {world} B200 '
                                    f'in chunks of <= {args.chunk} (BASELINE configs[2])', global_batch=args.global_batch,
                                    images_per_rank=r['images_per_rank'], chunk=r['chunk'], size=SIZE, cycle_align=2, mod_size=256,
                                    l2='every chunk (>= 100 MB of input, GBs of activations) exceeds the 126 MB L2',
                                    launch='CUDA-graph replay per chunk', parallelism=f'independent image shards x{world}, no collective'),
                        e2e=dict(value=r['e2e']['value'], unit=UNIT, ms_per_step=r['e2e']['ms_per_pass'],
                                 h2d_bytes_per_step=r['e2e']['h2d_bytes_per_pass_per_rank'] * world,
                                 d2h_bytes_per_step=r['e2e']['d2h_bytes_per_pass_per_rank'] * world) if 'e2e' in r else None,
                        gpu_launches=int(_lib.lib().ood_launch_count()))
            sys.stdout.flush()
            os.write(json_fd, (json.dumps(line) + '\n').encode())
        if world > 1:
            dist.destroy_process_group()
        return 0

    def step_resident():
        torch.manual_seed(1000 + rank)
        return net(x_dev)[0]

    # End-to-end serving loop through the public call net(x): every step copies its batch from pinned host memory and
    # returns its result to pinned host memory.  Copies run on their own streams so that H2D of step i+1 and D2H of
    # step i-1 overlap the compute of step i (double-buffered device input, full-duplex PCIe).
    h2d, d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    x_bufs = [torch.empty_like(x_dev) for _ in range(2)]
    out_stage = [torch.empty(B, 3, SIZE, SIZE, device=dev) for _ in range(2)]

    def run_e2e(n_steps):
        if pipe is not None:                         # graphs.PipelinedForward: the package's serving loop (no staging copies)
            for _ in range(n_steps):
                pipe.submit(x_host, out_host)
            pipe.synchronize()
            return out_host
        cur = torch.cuda.current_stream(dev)
        in_ready = [torch.cuda.Event() for _ in range(2)]
        done, d2h_done = [None, None], [None, None]
        with torch.cuda.stream(h2d):
            x_bufs[0].copy_(x_host, non_blocking=True)
            in_ready[0].record(h2d)
        for i in range(n_steps):
            k = i % 2
            cur.wait_event(in_ready[k])
            out = step_on(x_bufs[k])
            if graphed is not None:                  # the replay's output buffer is static: stage it so that the D2H copy
                if d2h_done[k] is not None:          # of step i can overlap the replay of step i+1
                    cur.wait_event(d2h_done[k])
                out_stage[k].copy_(out, non_blocking=True)
                out = out_stage[k]
            ev = torch.cuda.Event()
            ev.record(cur)
            done[k] = ev
            if i + 1 < n_steps:                      # prefetch the next batch into the other buffer
                with torch.cuda.stream(h2d):
                    if done[1 - k] is not None:
                        h2d.wait_event(done[1 - k])  # the step that last read that buffer has finished
                    x_bufs[1 - k].copy_(x_host, non_blocking=True)
                    in_ready[1 - k].record(h2d)
            with torch.cuda.stream(d2h):
                d2h.wait_event(ev)
                out.record_stream(d2h)
                out_host.copy_(out, non_blocking=True)
                d2h_done[k] = torch.cuda.Event()
                d2h_done[k].record(d2h)
        cur.wait_stream(d2h)
        return out

    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()

    if args.ncu:
        torch.cuda.profiler.start()
        step_resident()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return 0

    if args.profile and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            step_resident()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=40, max_name_column_width=70), file=sys.stderr)

    # The step is captured once into a CUDA graph and replayed (ood_gan_inversion_b200.graphs): ~550 launches per step are
    # otherwise issued from Python more slowly than a B200 retires the small ones.  --no-graph times eager launches.
    graphed, graph_note = None, 'eager launches (--no-graph)'
    if not args.no_graph:
        try:
            from ood_gan_inversion_b200.graphs import GraphedForward
            torch.manual_seed(1000 + rank)
            graphed = GraphedForward(lambda t: net(t)[0], x_dev, warmup=1)
            graph_note = 'CUDA-graph replay of net(x)'
        except Exception as e:                                        # capture is an optimisation, never a requirement
            graphed, graph_note = None, f'eager launches (graph capture failed: {type(e).__name__})'
            torch.cuda.synchronize()

    # end-to-end arm: two more captures of the same step, used round-robin by the package's serving loop
    pipe = None
    if graphed is not None:
        try:
            from ood_gan_inversion_b200.graphs import PipelinedForward
            torch.manual_seed(1000 + rank)
            pipe = PipelinedForward(lambda t: net(t)[0], x_dev, depth=2, warmup=1)
        except Exception as e:
            print(f'bench: PipelinedForward unavailable ({type(e).__name__}: {e}); e2e uses the staged loop', file=sys.stderr)
            pipe = None
            torch.cuda.synchronize()

    def step_on(t):
        if graphed is not None:
            return graphed(t)
        torch.manual_seed(1000 + rank)
        return net(t)[0]

    for _ in range(2):
        step_on(x_dev)
    barrier()

    # ---------------- timed region 1: inputs resident in HBM (inputs are 201 MB > 126 MB L2: no explicit flush) ----
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        e0.record()
        for _ in range(args.steps):
            out = step_on(x_dev)
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    if not torch.isfinite(out).all():
        raise RuntimeError('bench: non-finite output')

    # ---------------- per-kernel durations: the same step, launched eagerly with a CUDA event pair around every launch of
    # this library (a graph replay cannot carry per-launch events; kernel durations do not depend on how they were launched)
    launches0 = _lib.lib().ood_launch_count()
    K.profile_begin()
    prof_steps = min(args.steps, 3)
    for _ in range(prof_steps):
        step_resident()
    torch.cuda.synchronize()
    prof = K.profile_end()
    launches = (_lib.lib().ood_launch_count() - launches0) // prof_steps * args.steps

    # ---------------- timed region 2: end to end through the public call, host buffers ---------------------------
    run_e2e(2)
    barrier()
    e0.record()
    run_e2e(args.steps)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item())

    # ---------------- optional timed region 3: the same serving loop on the reference script's byte formats -----------------
    e2e_u8 = None
    if not args.no_u8_io and pipe is not None:
        try:
            from ood_gan_inversion_b200 import imgio
            from ood_gan_inversion_b200.graphs import PipelinedForward
            frames_dev = imgio.tensor2img(x_dev, min_max=(-1, 1))                      # uint8 BGR [B,1024,1024,3]
            frames_host = frames_dev.cpu().pin_memory()
            out8_host = torch.empty_like(frames_host).pin_memory()
            torch.manual_seed(1000 + rank)
            pipe8 = PipelinedForward(imgio.ByteServing(net), frames_dev, depth=2, warmup=1)
            for _ in range(2):
                pipe8.submit(frames_host, out8_host)
            pipe8.synchronize()
            barrier()
            e0.record()
            for _ in range(args.steps):
                pipe8.submit(frames_host, out8_host)
            pipe8.synchronize()
            e1.record()
            barrier()
            ms8 = reduce_max(e0.elapsed_time(e1))
            del pipe8
            e2e_u8 = dict(value=world * args.steps * B / (ms8 * 1e-3), unit=UNIT, ms_per_step=ms8 / args.steps,
                          h2d_bytes_per_step=frames_host.numel(), d2h_bytes_per_step=out8_host.numel(),
                          note='uint8 BGR frames in and out (run_ood_faceGAN_inversion.py:158-174 formats), converters inside the captured step')
        except Exception as exc:
            if world > 1:
                raise                                   # a rank that drops out of the collectives would hang the others
            e2e_u8 = dict(error=f'{type(exc).__name__}: {exc}'[:300])
            torch.cuda.synchronize()

    # ---------------- BASELINE configs[2]: one 256-image batch sharded over the ranks in chunks of <= 32 (every N) -------------
    config3 = None
    if not args.no_extra_legs:
        pipe = graphed = None                       # frees the three captured configs[1] graphs and their pools
        torch.cuda.empty_cache()
        config3 = sharded_batch_leg(net, 256, args.chunk, min(args.steps, 10), rank, world, dev, barrier, reduce_max)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    def ncu_traffic(name, summary='ncu_r01_summary.json'):
        """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the kernel family from the committed ncu --set full capture."""
        try:
            d = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', summary)))[name]
            unit = dict(byte=1.0, Kbyte=1e3, Mbyte=1e6, Gbyte=1e9)
            tot = 0.0
            for k in ('dram_read', 'dram_write'):
                v, u = d[k].split()
                tot += float(v) * unit[u]
            return tot
        except Exception:
            return None

    pk = peaks()
    conv = prof.get('conv3x3_tc', dict(ms=0.0, work=0.0, launches=0))
    blur = prof.get('blur_act', dict(ms=0.0, work=0.0, launches=0))
    conv_tf = conv['work'] / (conv['ms'] * 1e-3) / 1e12 if conv['ms'] > 0 else 0.0
    blur_gbs = blur['work'] / (blur['ms'] * 1e-3) / 1e9 if blur['ms'] > 0 else 0.0
    step_ms = ms_max / args.steps
    kern = {k: dict(ms_per_step=v['ms'] / prof_steps, launches_per_step=v['launches'] / prof_steps,
                    achieved=(v['work'] / (v['ms'] * 1e-3) / (1e12 if ('conv' in k and 'rows' not in k) else 1e9)) if v['ms'] > 0 else 0.0,
                    unit='TFLOP/s' if ('conv' in k and 'rows' not in k) else 'GB/s') for k, v in prof.items()}
    line = dict(metric=METRIC, value=world * args.steps * B / (ms_max * 1e-3), unit=UNIT, n_gpus=world, steps=args.steps,
                warmup=max(args.warmup, 3), ms_per_step=step_ms, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='bf16', data='synthetic',
                config=main_config(world, B, graph_note),
                clocks=clk.summary(),
                e2e=dict(value=world * args.steps * B / (ms_e2e * 1e-3), unit=UNIT, h2d_bytes_per_step=x_host.numel() * 4,
                         d2h_bytes_per_step=out_host.numel() * 4, ms_per_step=ms_e2e / args.steps),
                gpu_launches=int(launches),
                roofline=dict(kernel='conv_tc_kernel (tcgen05 implicit-GEMM convolutions of the generator and the AlignNet; the HBM-bound row kernels conv_rows / convt_rows are listed separately under kernels.conv3x3_rows)', bound='tensor', achieved=conv_tf,
                              peak=pk['tf_sus'], unit='TFLOP/s', frac=conv_tf / pk['tf_sus'],
                              traffic=ncu_traffic('conv256_pair', 'ncu_r02_summary.json') or ncu_traffic('conv256'),
                              traffic_note='DRAM bytes of one conv_tc_kernel<256,64,STATS,CTA pair> launch (AlignNet 1024->1024 ch at 64 px, batch 16: 1.24 TFLOP, 0.29 GB of activations + weights algorithmic), profiles/ncu_r02_conv256_pair_raw.csv (round-1 single-CTA capture: ncu_r01_conv256_raw.csv)',
                              peak_source=pk['src'] + ', sustained figure (kernel timed inside a long step)',
                              share_of_step=conv['ms'] / prof_steps / max(step_ms, 1e-9), launches_per_step=conv['launches'] / prof_steps,
                              timing='CUDA events around every launch in an eager pass of the same step, same process'),
                roofline_hbm=dict(kernel='blur_rows_kernel / blur_tma_kernel (FIR blur + demod + noise + bias + lrelu + next style)', bound='hbm',
                                  achieved=blur_gbs, peak=pk['hbm'], unit='GB/s', frac=blur_gbs / pk['hbm'], traffic=ncu_traffic('blurrows'),
                                  traffic_note='DRAM bytes of the 1024 px blur_rows_kernel launch (2.15 GB algorithmic), profiles/ncu_r01_blurrows_raw.csv',
                                  share_of_step=blur['ms'] / prof_steps / max(step_ms, 1e-9)),
                kernels=kern)
    if e2e_u8 is not None:
        line['e2e_u8'] = e2e_u8
    if config3 is not None:
        line['config3'] = config3
    if world == 1 and not args.no_extra_legs:
        # single-GPU context legs (rank 0 only; each guarded: a reported extra must not cost the line)
        state_dev = {k: v.detach() for k, v in net.state_dict().items()}
        for key, leg in (('parity', lambda: parity_leg(net, state_dev, x_dev, dev)),
                         ('gpu_reference', lambda: gpu_reference_leg(state_dev, dev)),
                         ('fp32_mode', lambda: fp32_mode_leg(net, state_dev, x_dev, dev)),
                         ('config4', lambda: inversion_leg(dev))):
            try:
                torch.cuda.empty_cache()
                line[key] = leg()
            except Exception as exc:
                line[key] = dict(error=f'{type(exc).__name__}: {exc}'[:300])
                torch.cuda.synchronize()
        sg.set_precision('bf16')
        del state_dev
    if world == 1 and not args.no_cpu_baseline:
        # the reference's CPU path, timed on this box's host cores by a child process that sees no GPU (--impl reference)
        try:
            env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
            o = subprocess.run([sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--steps', '2', '--warmup', '1',
                                '--no-cpu-baseline'], capture_output=True, text=True, timeout=600, env=env)
            rl = json.loads([l for l in o.stdout.splitlines() if l.startswith('{')][-1])
            line['cpu_baseline'] = {k: rl['cpu_baseline'][k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
        except Exception as exc:
            sd = {k: v.detach().float().cpu() for k, v in net.state_dict().items()}
            r = cpu_reference(2, 1, 1, state=sd, prefer_reference=False)
            line['cpu_baseline'] = dict(value=r['value'], unit=UNIT, cores=r['cores'], kind=r['kind'], sample=r['sample'],
                                        note=f'child process failed ({type(exc).__name__}); oracle port in-process')
        try:
            line['cpu_baseline']['ops'] = cpu_op_baselines(dev)['ops']
        except Exception as exc:                                 # a reported baseline must not cost the bench line
            line['cpu_baseline']['ops_error'] = f'{type(exc).__name__}: {exc}'[:300]
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + '\n').encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
