"""CUDA-graph capture of a shape-static piece of the operator graph.

One forward + vjp of a FastPM model on ONE GPU enqueues ~700 kernels from ~120 tape nodes and
never waits for the device: particle counts, mesh sizes and therefore every launch
configuration are fixed by the shapes, the cell sort and everything that depends on the values
happens on the device.  Such a call can be captured once into a CUDA graph and replayed with new
input VALUES (same buffers), which removes the Python VM, the ctypes calls and the launch gaps
from the step: the replay is one ``cudaGraphLaunch``.

Not applicable where the host reads device results inside the call: ``decompose`` across ranks
(the per-peer counts size the receive buffers), scalar reductions (``cnorm``, ``cdot``).

    step = GraphedCall(lambda: model.compute_with_vjp(init=dict(x=wn), v=dict(_dx=cot)))
    wn.t.copy_(new_values)        # inputs are the SAME device buffers the capture saw
    (dx,), (g,) = step()          # replay; outputs are the buffers of the captured call
"""
import torch

from .pm import runtime


class GraphedCall(object):
    def __init__(self, fn, warmup=2, release_memory=False):
        rt = runtime()
        self.fn = fn
        self.stream = torch.cuda.Stream(device=rt.device)
        cur = torch.cuda.current_stream(rt.device)
        self.stream.wait_stream(cur)
        # warm-up ON the capture stream: cuFFT plans, the context workspace and torch's
        # allocator reach their steady state, and the context already follows this stream when
        # the capture begins (switching streams synchronises, which a capture forbids)
        with torch.cuda.stream(self.stream):
            for _ in range(max(1, warmup)):
                fn()
        self.stream.synchronize()
        if release_memory:
            # the capture allocates from a private pool: hand the blocks the warm-up cached back
            # to the driver first, or a tape-sized problem (100 GB at 512^3) is resident twice
            import gc
            gc.collect()
            torch.cuda.empty_cache()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            self.outputs = fn()
        cur.wait_stream(self.stream)

    def __call__(self):
        """replay on torch's current stream; returns the captured call's output objects"""
        self.graph.replay()
        return self.outputs
