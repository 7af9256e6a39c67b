"""CUDA-graph replay of a fixed-shape forward pass.

The inversion forward is ~550 kernel launches per batch (this library's kernels, cuDNN / cuBLAS for the few remaining
library convolutions, the RNG for the injected noise).  At batch 16 on a B200 the launches of the small layers (4-64 px
generator levels, the encoder's tail) are issued more slowly from Python than the GPU retires them; capturing the step
once and replaying it removes that host-side bound.  Everything on the path is capture-safe: kernels are launched on the
current (capturing) stream, TMA descriptors travel by value as kernel parameters, workspaces come from the PyTorch
caching allocator (graph-private pool), there are no host synchronisations, and `torch.randn` uses the graph-safe
Philox offset of the default CUDA generator (every replay draws fresh noise).
"""
import torch


class GraphedForward:
    """graphed = GraphedForward(net, example_input); out = graphed(x)

    `x` must have the example's shape / dtype / device; the returned tensors are static buffers that the next call
    overwrites (copy them out -- or to the host -- before calling again)."""

    def __init__(self, fn, example, warmup=3, **kwargs):
        if not example.is_cuda:
            raise RuntimeError('ood_gan_inversion_b200 is CUDA-only')
        self.fn, self.kwargs = fn, kwargs
        self.static_in = example.clone()
        side = torch.cuda.Stream(example.device)
        side.wait_stream(torch.cuda.current_stream(example.device))
        with torch.cuda.stream(side), torch.no_grad():        # lazy initialisation (packed weights, smem attributes) happens here
            for _ in range(max(1, warmup)):
                fn(self.static_in, **kwargs)
        torch.cuda.current_stream(example.device).wait_stream(side)
        torch.cuda.synchronize(example.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = fn(self.static_in, **kwargs)

    def __call__(self, x):
        if x.shape != self.static_in.shape or x.dtype != self.static_in.dtype:
            raise ValueError(f'GraphedForward was captured for {tuple(self.static_in.shape)} {self.static_in.dtype}')
        if x.data_ptr() != self.static_in.data_ptr():
            self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out
