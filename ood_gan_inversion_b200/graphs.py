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


class PipelinedForward:
    """Serving loop over pinned host buffers (the host I/O either side of the path).

        pipe = PipelinedForward(lambda t: net(t)[0], example_batch)      # `depth` captured graphs, round-robin
        for x_host, out_host in requests:                                # pinned CPU tensors of the example's shape
            done = pipe.submit(x_host, out_host)                         # asynchronous: returns a CUDA event
        pipe.synchronize()                                               # or done.synchronize() per request

    Request i copies host -> device straight into the static input of graph i % depth on a copy stream, replays that graph
    on the current stream and copies its static output -> host on a second copy stream, so the H2D of request i+1 and the
    D2H of request i-1 overlap the replay of request i (full-duplex PCIe), with no device-to-device staging copies: an
    input buffer is rewritten only after the replay that read it, an output buffer only after its copy-out."""

    def __init__(self, fn, example, depth=2, warmup=1, **kwargs):
        if depth < 1:
            raise ValueError('PipelinedForward: depth must be >= 1')
        self.graphs = [GraphedForward(fn, example, warmup=warmup if i == 0 else 1, **kwargs) for i in range(depth)]
        dev = example.device
        self.device = dev
        self.h2d, self.d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.replayed = [None] * depth          # event: the last replay of graph k has finished (its input may be rewritten)
        self.copied_out = [None] * depth        # event: the last output of graph k has reached the host (may be overwritten)
        self.count = 0

    def submit(self, x_host, out_host):
        g = self.graphs[self.count % len(self.graphs)]
        k = self.count % len(self.graphs)
        if x_host.shape != g.static_in.shape or x_host.dtype != g.static_in.dtype:
            raise ValueError(f'PipelinedForward was captured for {tuple(g.static_in.shape)} {g.static_in.dtype}')
        if not isinstance(g.static_out, torch.Tensor):
            raise RuntimeError('PipelinedForward: the captured function must return one tensor')
        if out_host.shape != g.static_out.shape or out_host.dtype != g.static_out.dtype:
            raise ValueError(f'PipelinedForward returns {tuple(g.static_out.shape)} {g.static_out.dtype}')
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.h2d):
            if self.replayed[k] is not None:
                self.h2d.wait_event(self.replayed[k])
            g.static_in.copy_(x_host, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.h2d)
        cur.wait_event(ready)
        if self.copied_out[k] is not None:
            cur.wait_event(self.copied_out[k])
        g.graph.replay()
        fin = torch.cuda.Event()
        fin.record(cur)
        self.replayed[k] = fin
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(fin)
            out_host.copy_(g.static_out, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.d2h)
        self.copied_out[k] = done
        self.count += 1
        return done

    def synchronize(self):
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.d2h)
        cur.wait_stream(self.h2d)
        cur.synchronize()
