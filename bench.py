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


def cpu_reference(steps, warmup, images_per_step=1, state=None):
    """The reference's arithmetic for the path (oracle port: plain PyTorch fp32, native upfirdn2d / fused_act branch) on
    the host cores.  A step = `images_per_step` images of the same 1024px pipeline (bounded sample)."""
    import torch
    from oracle import ood as oood
    torch.set_num_threads(os.cpu_count() or 1)
    torch.set_grad_enabled(False)
    sd = state if state is not None else oood.synthetic_ood_state(SIZE, seed=0)
    from ood_gan_inversion_b200.synth import synthetic_faces
    x = synthetic_faces(images_per_step, SIZE)
    times = []
    for i in range(warmup + steps):
        torch.manual_seed(123)
        t0 = time.perf_counter()
        oood.ood_forward(sd, x, size=SIZE, strict_rng=False)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return dict(value=images_per_step * len(times) / total, ms_per_step=1e3 * total / len(times), cores=torch.get_num_threads(),
                sample=f'{len(times)} step(s) x {images_per_step} image(s) of the 1024px pipeline, fp32, after {warmup} warm-up')


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
    ap.add_argument('--u8-io', action='store_true', help='also time the byte-format serving loop (imgio.ByteServing: uint8 BGR frames '
                    'across PCIe both ways, 3 instead of 12 bytes per pixel) and report it as "e2e_u8"; N=1 only')
    ap.add_argument('--no-graph', action='store_true', help='time eager launches instead of a CUDA-graph replay of the step')
    ap.add_argument('--profile', action='store_true', help='print a torch.profiler kernel table for one step (not a bench value)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))

    if args.impl == 'reference':
        if rank != 0:
            return 0
        r = cpu_reference(max(1, args.steps), max(0, min(args.warmup, 1)))
        line = dict(metric=METRIC, value=r['value'], unit=UNIT, impl='reference', n_gpus=args.gpus, steps=args.steps,
                    warmup=min(args.warmup, 1), ms_per_step=r['ms_per_step'], higher_is_better=True, scaling='weak',
                    vs_baseline=None, dtype='f32', data='synthetic',
                    config=dict(workload=WORKLOAD, note='reference arm: the reference\'s own CPU arithmetic (native upfirdn2d / '
                                'fused_act branch, oracle port) on the host cores, 1 image per step'),
                    cpu_baseline=dict(value=r['value'], unit=UNIT, cores=r['cores'], kind='port', sample=r['sample']),
                    e2e=dict(value=r['value'], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
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
    torch.backends.cudnn.benchmark = True
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
    if args.u8_io and world == 1 and pipe is not None:
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
            ms8 = e0.elapsed_time(e1)
            e2e_u8 = dict(value=args.steps * B / (ms8 * 1e-3), unit=UNIT, ms_per_step=ms8 / args.steps,
                          h2d_bytes_per_step=frames_host.numel(), d2h_bytes_per_step=out8_host.numel(),
                          note='uint8 BGR frames in and out (run_ood_faceGAN_inversion.py:158-174 formats), converters inside the captured step')
        except Exception as exc:
            e2e_u8 = dict(error=f'{type(exc).__name__}: {exc}'[:300])
            torch.cuda.synchronize()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    def ncu_traffic(name):
        """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the kernel family from the committed ncu --set full capture."""
        try:
            d = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', 'ncu_r01_summary.json')))[name]
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
                    achieved=(v['work'] / (v['ms'] * 1e-3) / (1e12 if 'conv' in k else 1e9)) if v['ms'] > 0 else 0.0,
                    unit='TFLOP/s' if 'conv' in k else 'GB/s') for k, v in prof.items()}
    line = dict(metric=METRIC, value=world * args.steps * B / (ms_max * 1e-3), unit=UNIT, n_gpus=world, steps=args.steps,
                warmup=max(args.warmup, 3), ms_per_step=step_ms, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='bf16', data='synthetic',
                config=dict(workload=WORKLOAD, batch_per_gpu=B, global_batch=B * world, size=SIZE, cycle_align=2, mod_size=256,
                            l2='inputs (201 MB/step) exceed the 126 MB L2', launch=graph_note, parallelism=f'independent image shards x{world}, no collective',
                            weights='random-init (reference init + non-zero noise weights), synthetic smooth faces'),
                clocks=clk.summary(),
                e2e=dict(value=world * args.steps * B / (ms_e2e * 1e-3), unit=UNIT, h2d_bytes_per_step=x_host.numel() * 4,
                         d2h_bytes_per_step=out_host.numel() * 4, ms_per_step=ms_e2e / args.steps),
                gpu_launches=int(launches),
                roofline=dict(kernel='conv_tc_kernel (tcgen05 implicit-GEMM 3x3 modulated conv)', bound='tensor', achieved=conv_tf,
                              peak=pk['tf_sus'], unit='TFLOP/s', frac=conv_tf / pk['tf_sus'], traffic=ncu_traffic('conv256'),
                              traffic_note='DRAM bytes of one conv_tc_kernel<256,64,STATS> launch (AlignNet 1024->1024 ch at 64 px: 1.24 TFLOP, 0.29 GB of activations + weights algorithmic; L2 hit 96 %), profiles/ncu_r01_conv256_raw.csv',
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
    if world == 1 and not args.no_cpu_baseline:
        sd = {k: v.detach().float().cpu() for k, v in net.state_dict().items()}
        r = cpu_reference(2, 1, 1, state=sd)
        line['cpu_baseline'] = dict(value=r['value'], unit=UNIT, cores=r['cores'], kind='port', sample=r['sample'])
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
