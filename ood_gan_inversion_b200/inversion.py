"""Optimisation-based latent inversion (BASELINE config 4): Adam on W+ latents through the differentiable synthesis.

Reference behaviour: the generator is frozen (fix list options/train/E4E_Face.yml:123-125) and the leaves are either one
W+ code per image (latent optimisation: autograd through Generator.forward, src/ops/StyleGAN/model.py:483-585, pixel loss
`pix_opt`, E4E_Face.yml:152-154) or ONE offset shared by every image, the arch's `delta_latent[1, n_latent, 512]`
(src/archs/OOD_faceGAN_e4e_arch.py:126-129, `lats = w + avg + delta`, :283-286).

Multi-GPU: images shard across ranks.  Per-image codes need no communication at all.  The shared
offset is the path's one real exchange step: every rank back-propagates its own images, the [1, n_latent, 512] gradient
(36 KB) is all-reduced -- NCCL over NVLink on the GPU box, gloo in the CPU tests -- and every rank applies the identical Adam
update, so the offsets stay bit-identical across ranks without a broadcast.

The synthesis itself is `Generator.forward` of this package: with a latent that requires grad it runs the hand-written
backward of synthesis_grad.SynthesisFn (CUDA only; there is no CPU fallback).  `synthesize` can be replaced by any
differentiable callable latent -> image, which is how the host-side logic is tested without a GPU.
"""
import os

import torch
import torch.distributed as dist

_GRAPH = os.environ.get('OOD_INVERSION_GRAPH', '0') != '0'      # replay the Adam step as a CUDA graph (see LatentInverter.run); off: measured, no gain
_GRAPH_EAGER_STEPS = 3                                           # eager steps before the capture (lazy initialisation, allocator steady state)


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group)
    return 1


def generator_synthesizer(gen, noise=None):
    """latent [B, n_latent, D] -> image through `gen` (W+ input, fixed noise: the registered buffers unless `noise` is given)."""
    def synthesize(latent):
        return gen(latent, input_is_tensor=True, input_is_latent=True, randomize_noise=False, noise=noise)[0]
    return synthesize


def blended_synthesizer(gen, fields, x, noise=None):
    """BASELINE config 4 with the blend in the loop ("generator (+ blend optional)"): latent -> A * x + gen(latent) * (1 - A) with the
    invertibility masks `fields` (the arch's `aligns[1..4]`, fp32 [B,3,r,r], held fixed) and the input batch `x`
    (OOD_faceGAN_e4e_arch.py:315-347).  The blend runs the inference kernel forward and ood_mask_blend_bwd backward
    (samm_grad.mask_blend), so the OOD region of the target does not pull on the latent."""
    from . import samm_grad
    fields = [f.detach() for f in fields]
    x = x.detach()

    def synthesize(latent):
        image = gen(latent, input_is_tensor=True, input_is_latent=True, randomize_noise=False, noise=noise)[0]
        return samm_grad.mask_blend(fields, x, image)[0]
    return synthesize


def mse_loss(image, target):
    """`pix_opt` of the reference (L2, reduction mean; E4E_Face.yml:152-154)."""
    return torch.nn.functional.mse_loss(image, target)


class LatentInverter:
    """inv = LatentInverter(synthesize, lr=0.01, shared_delta=False); latent, losses = inv.run(target, base_latent, steps)

    shared_delta=False: one code per image; leaves = `base_latent.clone()` [B, n_latent, D]; no collective.
    shared_delta=True : leaves = delta [1, n_latent, D] (zeros); image = synthesize(base_latent + delta); the loss is the mean
                        over the LOCAL images, its gradient is averaged over the ranks of `group` (every rank must hold the
                        same number of images, as sharding.shard_bounds gives for a divisible batch).
    """

    def __init__(self, synthesize, lr=0.01, shared_delta=False, loss=mse_loss, group=None, betas=(0.9, 0.999), graph=None):
        self.synthesize, self.lr, self.shared_delta, self.loss, self.group, self.betas = synthesize, lr, shared_delta, loss, group, betas
        self.graph = _GRAPH if graph is None else bool(graph)

    def _check_equal_shards(self, n_local, device):
        if _world(self.group) == 1:
            return
        t = torch.tensor([n_local, -n_local], device=device, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        if int(t[0]) != -int(t[1]):
            raise ValueError('LatentInverter(shared_delta=True): every rank must hold the same number of images '
                             f'(this rank {n_local}, max {int(t[0])}, min {-int(t[1])})')

    def run(self, target, base_latent, steps, callback=None):
        """target [B,3,H,W]; base_latent [B, n_latent, D] (e.g. the encoder's codes or the repeated average latent).
        Returns (latent [B, n_latent, D] = the optimised codes, losses: list of per-step python floats of the LOCAL loss)."""
        if base_latent.shape[0] != target.shape[0]:
            raise ValueError('LatentInverter: one base latent per target image')
        base = base_latent.detach()
        if self.shared_delta:
            self._check_equal_shards(target.shape[0], target.device)
            leaf = torch.zeros((1,) + tuple(base.shape[1:]), device=base.device, dtype=base.dtype, requires_grad=True)
        else:
            leaf = base.clone().requires_grad_(True)
        world = _world(self.group)
        # One Adam step is a fixed sequence of ~170 launches (forward, hand-written backward, the optimizer's foreach kernels); with graph=True (or
        # OOD_INVERSION_GRAPH=1) it is captured into a CUDA graph after a few eager steps and replayed.  Measured at batch 32, 1024 px: 23.8-23.9 ms
        # per replayed step against 23.7-24.2 ms eager (CUDA events around the replays) -- the eager loop is not launch-bound, so this is off by
        # default.  Never with a per-step callback or when the step contains a collective (shared offset across ranks).
        use_graph = self.graph and leaf.is_cuda and callback is None and not (self.shared_delta and world > 1) and steps > _GRAPH_EAGER_STEPS + 1
        opt = torch.optim.Adam([leaf], lr=self.lr, betas=self.betas, capturable=bool(use_graph))
        losses = []

        def step():
            latent = base + leaf if self.shared_delta else leaf
            loss = self.loss(self.synthesize(latent), target)
            loss.backward()
            if self.shared_delta and world > 1:
                dist.all_reduce(leaf.grad, op=dist.ReduceOp.SUM, group=self.group)      # 36 KB at [1,18,512] fp32
                leaf.grad.div_(world)
            opt.step()
            return loss.detach()

        n_eager = _GRAPH_EAGER_STEPS if use_graph else steps
        for i in range(n_eager):
            opt.zero_grad(set_to_none=True)
            loss = step()
            losses.append(loss)
            if callback is not None:
                callback(i, loss, leaf.detach())
        if use_graph:
            dev = leaf.device
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            opt.zero_grad(set_to_none=True)                       # the captured backward allocates leaf.grad inside the graph's pool
            with torch.cuda.graph(graph):
                static_loss = step()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n_eager, steps):
                graph.replay()
                losses.append(static_loss.clone())
            e1.record()
            torch.cuda.current_stream(dev).synchronize()
            # device time of the replayed steps of this run (the eager steps and the capture before them are a fixed cost of a run)
            self.timing = dict(eager_steps=n_eager, replay_steps=steps - n_eager, replay_ms=e0.elapsed_time(e1))
            del graph
        else:
            self.timing = dict(eager_steps=steps, replay_steps=0, replay_ms=0.0)
        self.delta = leaf.detach() if self.shared_delta else None         # the shared offset itself (bit-identical on every rank)
        out = (base + leaf.detach()) if self.shared_delta else leaf.detach()
        return out, [float(l) for l in losses]


def invert(gen, target, base_latent, steps=500, lr=0.01, shared_delta=False, noise=None, group=None):
    """Config 4 in one call: `steps` Adam steps of pixel-MSE inversion of `target` through `gen` (frozen)."""
    params = list(gen.parameters())
    was = [p.requires_grad for p in params]
    for p in params:
        p.requires_grad_(False)                    # the generator is frozen for the run (options/train/E4E_Face.yml:123-125) ...
    try:
        return LatentInverter(generator_synthesizer(gen, noise), lr=lr, shared_delta=shared_delta, group=group).run(target, base_latent, steps)
    finally:
        for p, r in zip(params, was):
            p.requires_grad_(r)                    # ... and handed back as it was
