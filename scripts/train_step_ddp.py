"""Training step of the trainable parts (`modulation` + `feats_conv`, options/train/E4E_Face.yml:123-125) under data parallelism:
one process per GPU (torchrun), images sharded, gradients averaged by training.GradAllReduce (bucketed NCCL all-reduce launched from
autograd hooks, overlapping the rest of the backward).  Prints per-step time with the overlapped exchange and with one blocking
exchange after the backward, and checks that every rank ends with identical parameters.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/train_step_ddp.py --batch 2 --steps 4
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=2)
ap.add_argument('--steps', type=int, default=4)
ap.add_argument('--mod-size', type=int, default=256)
args = ap.parse_args()
rank, local, world = int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
from ood_gan_inversion_b200 import stylegan as sg
from ood_gan_inversion_b200.arch import ood_faceGAN_e4e
from ood_gan_inversion_b200.synth import synthetic_faces, synthetic_init
from ood_gan_inversion_b200.training import GradAllReduce, apply_fix_list, generator_step
sg.set_precision('bf16')
torch.manual_seed(0)
net = synthetic_init(ood_faceGAN_e4e(out_size=1024, style_dim=512, encoder='E4E', enable_modulation=True, warp_scale=0.08, cycle_align=2,
                                     blend_with_gen=True, ModSize=args.mod_size, eval_path_length=False), seed=0).to(dev).eval()
trainable = apply_fix_list(net)
params = [p for _, p in trainable]
nbytes = sum(p.numel() * 4 for p in params)
x = synthetic_faces(args.batch, 1024, seed=10 + rank, device=dev)
target = (0.8 * x).detach()
opt = torch.optim.Adam(params, lr=1e-4)
res = dict(world=world, batch_per_gpu=args.batch, trainable_tensors=len(params), gradient_mbytes=nbytes / 1e6)
for mode in ('overlapped', 'blocking'):
    sync = GradAllReduce(params, bucket_mb=25.0)
    if mode == 'blocking':
        sync.remove()                                   # no hooks: finish() launches every bucket after the backward
    for _ in range(2):
        generator_step(net, x, target, opt, sync=sync)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = generator_step(net, x, target, opt, sync=sync)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[f'ms_per_step_{mode}'] = float(t)
    sync.remove()
chk = torch.stack([p.detach().double().sum() for p in params]).sum().reshape(1)
if world > 1:
    both = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(both, chk)
    res['parameters_identical_across_ranks'] = bool(all(torch.equal(b, both[0]) for b in both))
res['loss'] = float(loss)
res['peak_mem_gib'] = torch.cuda.max_memory_allocated() / 2 ** 30
if rank == 0:
    print(json.dumps(res))
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(res, open(f'gpurun_out/train_step_ddp_{world}gpu.json', 'w'), indent=1)
if world > 1:
    dist.destroy_process_group()
