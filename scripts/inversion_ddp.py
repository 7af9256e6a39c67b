"""Shared-offset latent inversion across ranks (the path's one exchange step): every rank inverts its
own slice of the images through the 1024 px generator with ONE shared delta_latent [1,18,512]; the 36 KB gradient is
all-reduced over NCCL every Adam step.  Launch: torchrun --nproc-per-node N scripts/inversion_ddp.py [--batch 8 --steps 20].
Prints one JSON line from rank 0: ms per step (max over ranks, CUDA events), the loss curve ends, and whether the offsets of
all ranks are bit-identical."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=8, help='images per rank')
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--size', type=int, default=1024)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    from ood_gan_inversion_b200 import stylegan as sg
    from ood_gan_inversion_b200.inversion import LatentInverter, generator_synthesizer
    from ood_gan_inversion_b200.synth import synthetic_faces, synthetic_init
    sg.set_precision('bf16')
    torch.manual_seed(0)
    gen = synthetic_init(sg.Generator(args.size, 512, 8), seed=0).to(dev)
    for p in gen.parameters():
        p.requires_grad_(False)
    target = synthetic_faces(args.batch, args.size, seed=3 + rank, device=dev)       # a different slice of images per rank
    base = torch.zeros(args.batch, gen.n_latent, 512, device=dev)
    inv = LatentInverter(generator_synthesizer(gen), lr=0.01, shared_delta=True)
    inv.run(target, base, 2)                                                         # warm-up (packs weights, sizes the allocator)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _, losses = inv.run(target, base, args.steps)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    same = True
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gathered = [torch.empty_like(inv.delta) for _ in range(world)]
        dist.all_gather(gathered, inv.delta.contiguous())
        same = all(torch.equal(g, gathered[0]) for g in gathered)
    if rank == 0:
        print(json.dumps(dict(what='shared delta_latent inversion, gradient all-reduce over NCCL', n_gpus=world, batch_per_gpu=args.batch,
                              size=args.size, steps=args.steps, ms_per_step=float(t.item()),
                              image_steps_per_s=world * args.batch * 1e3 / float(t.item()), loss_first=losses[0], loss_last=losses[-1],
                              allreduce_bytes_per_step=int(inv.delta.numel() * 4), offsets_identical_across_ranks=bool(same))), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
