"""Training side of the path: which parameters train, one generator step, and the data-parallel
gradient exchange.

Reference: src/models/OOD_faceGAN_model.py:663-789 (G step: forward, losses, `l_total.backward()`, optimizer step),
options/train/E4E_Face.yml:123-125 (`fix_and_grad.fix: ['generator', 'avg_latent', 'encoder']` -- what trains is `modulation`, the four
AlignNets, and `feats_conv`, ~44 M parameters / 176 MB of fp32 gradients) and BasicSR base_model.py:87-101 (DistributedDataParallel).

Images shard across ranks with no data-path collective (sharding.py); the one exchange of a training step is the gradient all-reduce.
GradAllReduce does what DDP's reducer does, without wrapping the module: parameters are grouped into fixed-size buckets in reverse
registration order (the order autograd produces their gradients), a post-accumulate hook launches an asynchronous all-reduce (NCCL over
NVLink on the box, gloo in the CPU tests) as soon as a bucket is complete -- so the exchange overlaps the rest of the backward (the
coarser AlignNet levels and the generator's data-gradient convolutions) -- and `finish()` waits, averages and scatters back.  torch's
own DistributedDataParallel works on this package's modules as well (every gradient is produced by autograd Functions).
"""
import torch
import torch.distributed as dist

FIX_DEFAULT = ('generator', 'avg_latent', 'encoder')


def apply_fix_list(net, fix=FIX_DEFAULT, grad=()):
    """The reference's `fix_and_grad` (OOD_faceGAN_model.py:545-576): parameters whose name contains an entry of `fix` are frozen, entries
    of `grad` are re-enabled afterwards.  Returns the list of trainable (name, parameter) pairs."""
    for name, p in net.named_parameters():
        frozen = any(f in name for f in fix)
        if any(g in name for g in grad):
            frozen = False
        p.requires_grad_(not frozen and p.is_floating_point())
    return [(n, p) for n, p in net.named_parameters() if p.requires_grad]


class GradAllReduce:
    """Bucketed, overlapped gradient averaging over `group`.

        sync = GradAllReduce([p for _, p in trainable], bucket_mb=25)
        loss.backward()          # hooks launch one all-reduce per completed bucket
        sync.finish()            # wait + average + write back into .grad; call before optimizer.step()

    Parameters that received no gradient in a step (a progressive stage that skips a level) contribute zeros, like DDP with
    find_unused_parameters.  World size 1: a no-op."""

    def __init__(self, params, bucket_mb=25.0, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.buckets, cur, size = [], [], 0
        limit = int(bucket_mb * 2 ** 20)
        for p in reversed(self.params):                       # gradients arrive roughly in reverse registration order
            nbytes = p.numel() * p.element_size()
            if cur and size + nbytes > limit:
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += nbytes
        if cur:
            self.buckets.append(cur)
        self.bucket_of = {id(p): bi for bi, b in enumerate(self.buckets) for p in b}
        self._reset()
        self.hooks = []
        if self.world > 1:
            for p in self.params:
                self.hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _reset(self):
        self.pending = [len(b) for b in self.buckets]
        self.inflight = {}

    def _launch(self, bi):
        ps = self.buckets[bi]
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in ps])
        self.inflight[bi] = (flat, dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _on_grad(self, p):
        bi = self.bucket_of[id(p)]
        self.pending[bi] -= 1
        if self.pending[bi] == 0:
            self._launch(bi)

    def finish(self):
        if self.world == 1:
            return
        for bi in range(len(self.buckets)):
            if bi not in self.inflight:                       # a bucket with parameters that got no gradient this step
                self._launch(bi)
        for bi, (flat, work) in self.inflight.items():
            work.wait()
            flat.div_(self.world)
            off = 0
            for p in self.buckets[bi]:
                n = p.numel()
                g = flat[off:off + n].view_as(p)
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
                off += n
        self._reset()

    def remove(self):
        for h in self.hooks:
            h.remove()
        self.hooks = []


def generator_step(net, x, target, optimizer, loss_fn=None, sync=None):
    """One optimisation step of the trainable parts (modulation / feats_conv) with a pixel loss on the blended output -- the skeleton of
    OOD_faceGAN_model.py:663-789 (the reference adds perceptual / mask / adversarial terms, which are outside the hot path).
    Returns the detached loss."""
    loss_fn = loss_fn or torch.nn.functional.mse_loss
    optimizer.zero_grad(set_to_none=True)
    out, _ = net(x)
    loss = loss_fn(out, target)
    loss.backward()
    if sync is not None:
        sync.finish()
    optimizer.step()
    return loss.detach()
