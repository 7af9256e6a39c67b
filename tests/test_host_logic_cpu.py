"""Host-side logic that needs no GPU: module trees / state-dict keys, pad arithmetic, conditioning schedule, loud
CPU-tensor errors, and the N>1 sharding + max-over-ranks reduction over gloo (world_size 2)."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

from oracle import ood as oood, stylegan as ostyle


def test_generator_state_dict_keys_match_reference_layout():
    from ood_gan_inversion_b200.stylegan import Generator
    for size in (16, 256, 1024):
        sd = ostyle.synthetic_generator_state(size, seed=1)     # key names pinned against the reference by make_golden.py
        g = Generator(size, 512, 8)
        assert set(g.state_dict().keys()) == set(sd.keys())
        g.load_state_dict(sd, strict=True)
    assert len(Generator(1024, 512, 8).state_dict()) == 171
    assert Generator(1024, 512, 8).n_latent == 18


def test_full_arch_state_dict_and_schedule():
    from ood_gan_inversion_b200.arch import ood_faceGAN_e4e
    net = ood_faceGAN_e4e(out_size=1024, warp_scale=0.08, cycle_align=2, ModSize=256)
    sd = oood.synthetic_ood_state(1024, seed=0)
    net.load_state_dict(sd, strict=True)
    assert len(net.state_dict()) == 882
    feats = [torch.zeros(1, 1, r, r) for r in (256, 128, 64, 32)]
    conds = net.feats2condition(feats)
    assert len(conds) == 4 and [(2 * (k + 2)) + 1 for k in range(len(conds))] == [5, 7, 9, 11]
    net.ModSize = 64
    assert len(net.feats2condition(feats)) == 2
    n_params = sum(p.numel() for p in net.parameters())
    assert abs(n_params - 341.49e6) < 0.05e6


def test_pad_arithmetic_and_kernels():
    from ood_gan_inversion_b200.stylegan import Blur, Downsample, ModulatedConv2d, Upsample, make_kernel
    k = make_kernel([1, 3, 3, 1])
    assert k.shape == (4, 4) and abs(float(k.sum()) - 1) < 1e-6
    assert Upsample([1, 3, 3, 1]).pad == (2, 1) and float(Upsample([1, 3, 3, 1]).kernel.sum()) == pytest.approx(4.0)
    assert Downsample([1, 3, 3, 1]).pad == (1, 1)
    assert ModulatedConv2d(8, 8, 3, 8, upsample=True).blur.pad == (1, 1)
    assert ModulatedConv2d(8, 8, 3, 8, downsample=True).blur.pad == (2, 2)
    assert Blur([1, 3, 3, 1], (1, 1), upsample_factor=2).taps == pytest.approx([0.25, 0.75, 0.75, 0.25])
    assert ostyle.up_blur_pads() == (1, 1) and ostyle.skip_up_pads() == (2, 1)


def test_cpu_tensors_raise():
    from ood_gan_inversion_b200.op import fused_leaky_relu, upfirdn2d
    from ood_gan_inversion_b200.stylegan import Generator
    with pytest.raises(RuntimeError, match='CUDA-only'):
        upfirdn2d(torch.zeros(1, 1, 4, 4), torch.ones(4, 4))
    with pytest.raises(RuntimeError, match='CUDA-only'):
        fused_leaky_relu(torch.zeros(2, 4), torch.zeros(4))
    with pytest.raises(RuntimeError, match='CUDA-only'):
        Generator(16, 512, 2)(torch.zeros(1, 6, 512), input_is_tensor=True, input_is_latent=True)


def test_shard_bounds_cover_the_batch():
    from ood_gan_inversion_b200.sharding import shard_bounds
    for total, world in [(256, 8), (256, 4), (16, 3), (5, 8), (0, 2)]:
        spans = [shard_bounds(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from ood_gan_inversion_b200.sharding import aggregate_throughput, max_over_ranks, shard_bounds
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    lo, hi = shard_bounds(33, world, rank)
    ms = 100.0 * (rank + 1)                     # rank 1 is the slow one
    slowest = max_over_ranks(ms)
    thr = aggregate_throughput(hi - lo, ms)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, slowest, thr, gathered))


def test_two_rank_gloo_sharding_and_timing():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, slowest, thr, gathered in res:
        assert slowest == pytest.approx(200.0)                       # max over ranks, not the local time
        assert thr == pytest.approx(33 / 0.2)                        # all units / slowest rank
        assert gathered == [(0, 17), (17, 33)]


# ----------------------------------------------------------------------------------------------- inversion (config 4) host logic
def _toy_problem(n, seed=0):
    """A differentiable stand-in for the synthesis (the real one is CUDA-only): image = tanh(latent . A)."""
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(4 * 8, 3 * 6 * 6, generator=g) * 0.3
    base = torch.randn(n, 4, 8, generator=g) * 0.1
    target = torch.tanh((torch.randn(n, 4, 8, generator=g)).reshape(n, -1) @ a).reshape(n, 3, 6, 6)
    synth = lambda lat: torch.tanh(lat.reshape(lat.shape[0], -1) @ a).reshape(-1, 3, 6, 6)
    return synth, base, target


def test_latent_inverter_single_process_modes():
    from ood_gan_inversion_b200.inversion import LatentInverter
    synth, base, target = _toy_problem(6)
    lat, losses = LatentInverter(synth, lr=0.05).run(target, base, 40)
    assert lat.shape == base.shape and losses[-1] < 0.5 * losses[0]
    lat_d, losses_d = LatentInverter(synth, lr=0.05, shared_delta=True).run(target, base, 40)
    delta = lat_d - base
    assert torch.allclose(delta[0], delta[-1], atol=1e-6) and losses_d[-1] < losses_d[0]     # ONE offset for every image
    with pytest.raises(ValueError):
        LatentInverter(synth).run(target, base[:2], 1)


def _inversion_worker(rank, world, port, q, uneven):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from ood_gan_inversion_b200.inversion import LatentInverter
    from ood_gan_inversion_b200.sharding import shard_bounds
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    synth, base, target = _toy_problem(8)
    lo, hi = shard_bounds(8, world, rank)
    if uneven and rank == 1:
        hi -= 1
    try:
        inv = LatentInverter(synth, lr=0.05, shared_delta=True)
        lat, losses = inv.run(target[lo:hi], base[lo:hi], 12)
        res = ('ok', inv.delta[0].tolist(), losses)      # plain lists: the worker exits right after the put
    except ValueError as e:
        res = ('error', str(e), None)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank,) + res)


def _run_two_ranks(uneven):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000 + (1 if uneven else 0)
    procs = [ctx.Process(target=_inversion_worker, args=(r, 2, port, q, uneven)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_shared_delta_inversion_allreduce_matches_single_process():
    """the shared delta_latent is the one exchange step of the path -- two gloo ranks on half the images each
    must walk the same Adam trajectory as one process on all of them, and stay identical to each other."""
    from ood_gan_inversion_b200.inversion import LatentInverter
    res = _run_two_ranks(uneven=False)
    assert [r[1] for r in res] == ['ok', 'ok']
    d0, d1 = torch.tensor(res[0][2]), torch.tensor(res[1][2])
    assert torch.equal(d0, d1)
    synth, base, target = _toy_problem(8)
    inv = LatentInverter(synth, lr=0.05, shared_delta=True)
    lat, losses = inv.run(target, base, 12)
    assert torch.allclose(inv.delta[0], d0, atol=1e-6)
    # the global loss is the mean of the two local ones (equal shard sizes)
    assert losses[-1] == pytest.approx(0.5 * (res[0][3][-1] + res[1][3][-1]), rel=1e-5)


def test_shared_delta_inversion_rejects_uneven_shards():
    res = _run_two_ranks(uneven=True)
    assert [r[1] for r in res] == ['error', 'error'] and 'same number of images' in res[0][2]


def _run_bench(*args, env=None):
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(root, 'bench.py'), *args], capture_output=True, text=True, env=e, timeout=600)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference`: one JSON line on rank 0 with the contract's keys (CPU arithmetic of the path, bounded sample,
    per-op baselines of SURVEY 8(d)); every other rank exits 0 without work or output."""
    import json
    r = _run_bench('--impl', 'reference', '--steps', '1', '--warmup', '0')
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == '1024px inversion images/sec' and d['unit'] == 'images/s'
    assert d['higher_is_better'] is True and d['vs_baseline'] is None and d['gpu_launches'] == 0
    assert d['value'] > 0 and d['e2e'] == dict(value=d['value'], unit='images/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    cb = d['cpu_baseline']
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    staged = os.path.isfile(os.path.join(root, 'baseline', '_ref', 'src', 'archs', 'OOD_faceGAN_e4e_arch.py'))
    # the unmodified reference when __graft_entry__.build() staged it (baseline/_ref, git-ignored), else the oracle port
    assert cb['kind'] == ('reference' if staged else 'port') and cb['cores'] >= 1 and cb['value'] == d['value'] and 'image' in cb['sample']
    assert d['config']['workload'].startswith('E4E encoder + StyleGAN2 1024px') and d['config']['global_batch'] == 16
    ops = {o['op']: o for o in cb['ops']}
    assert len(ops) == 5 and all(o['cpu_ms'] > 0 for o in ops.values())
    assert any('Generator(256)' in k for k in ops) and 'fused_leaky_relu' in ops
    other = _run_bench('--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0', env=dict(RANK='1', LOCAL_RANK='1', WORLD_SIZE='2'))
    assert other.returncode == 0 and other.stdout.strip() == ''


def test_bench_product_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip('CPU-only check')
    r = _run_bench('--steps', '1', '--warmup', '0', '--no-cpu-baseline')
    assert r.returncode != 0 and 'no CPU fallback' in r.stderr and '{' not in r.stdout


def _grad_allreduce_worker(rank, world, port, out):
    import torch.distributed as dist
    from ood_gan_inversion_b200.training import GradAllReduce
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(64, 300), torch.nn.ReLU(), torch.nn.Linear(300, 200), torch.nn.ReLU(), torch.nn.Linear(200, 8))
        unused = torch.nn.Parameter(torch.ones(5))                    # never receives a gradient
        params = list(net.parameters()) + [unused]
        sync = GradAllReduce(params, bucket_mb=0.1)                   # ~26k floats per bucket: several buckets
        assert len(sync.buckets) >= 3
        xs = torch.randn(8, 64, generator=torch.Generator().manual_seed(1))
        ys = torch.randn(8, 8, generator=torch.Generator().manual_seed(2))
        lo, hi = rank * 4, rank * 4 + 4
        for step in range(2):                                         # two steps: the hook state resets
            for p in params:
                p.grad = None
            torch.nn.functional.mse_loss(net(xs[lo:hi]), ys[lo:hi]).backward()
            sync.finish()
        got = [p.grad.clone() for p in params]
        for p in params:
            p.grad = None
        torch.nn.functional.mse_loss(net(xs), ys).backward()          # the full batch on one process
        ok = all(torch.allclose(g, p.grad if p.grad is not None else torch.zeros_like(p), rtol=1e-5, atol=1e-6) for g, p in zip(got, params))
        out.put((rank, ok, len(sync.buckets)))
    finally:
        dist.destroy_process_group()


def test_grad_allreduce_buckets_gloo_world2():
    """training.GradAllReduce (the training path's one collective): bucketed, hook-launched all-reduce over gloo, world size 2;
    the averaged shard gradients equal the full-batch gradient, unused parameters average to zero, state resets between steps."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_grad_allreduce_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res


def test_apply_fix_list_matches_the_reference_yml():
    """options/train/E4E_Face.yml:123-125: generator, avg_latent and encoder are frozen; modulation and feats_conv train."""
    from ood_gan_inversion_b200.training import apply_fix_list

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.generator = torch.nn.Linear(2, 2)
            self.encoder = torch.nn.Linear(2, 2)
            self.modulation = torch.nn.ModuleList([torch.nn.Linear(2, 2)])
            self.feats_conv = torch.nn.ModuleList([torch.nn.Conv2d(2, 2, 1)])
            self.avg_latent = torch.nn.Parameter(torch.zeros(1, 2))
            self.delta_latent = torch.nn.Parameter(torch.zeros(1, 18, 2))
    net = Tiny()
    names = [n for n, _ in apply_fix_list(net)]
    assert all(n.startswith(('modulation', 'feats_conv', 'delta_latent')) for n in names) and any(n.startswith('modulation') for n in names)
    assert not net.generator.weight.requires_grad and not net.avg_latent.requires_grad and net.feats_conv[0].bias.requires_grad
    names = [n for n, _ in apply_fix_list(net, grad=('encoder',))]
    assert any(n.startswith('encoder') for n in names)
