"""CUDA-graph replay of a whole training step (forward + CCC loss + backward + optimizer step).

The tape engine never synchronises with the host and takes every buffer from PyTorch's caching allocator,
so one step is capturable as a single CUDA graph: ~380 kernel launches (ctypes calls, tensor-map encodes,
Python tape closures) collapse into one `cudaGraphLaunch`, which is what keeps the GPU fed once a step is
only ~20 ms long -- and what matters at 8 GPUs, where every rank pays the host cost (SURVEY 7 "hard parts" #2).

The inputs live in caller-provided static device buffers; one graph is captured per buffer set (sharing one
memory pool) so the host->device copy of step i+1 can land in the other set while step i runs.
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import torch

from . import _lib as L


class GraphedStep:
    def __init__(self, step_fn: Callable[..., torch.Tensor], buffer_sets: Sequence[Sequence[torch.Tensor]], warmup: int = 3):
        """step_fn(*buffers) -> scalar loss tensor; it must do zero_grad / backward / optimizer.step itself."""
        assert len(buffer_sets) >= 1
        self.buffer_sets = [tuple(b) for b in buffer_sets]
        dev = self.buffer_sets[0][0].device
        try:    # the warm-up runs on a side stream: AccumulateGrad nodes created before it would warn about the stream switch
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        except AttributeError:
            pass
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                      # torch requires warm-up iterations on a side stream
            for i in range(max(warmup, 1)):
                step_fn(*self.buffer_sets[i % len(self.buffer_sets)])
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graphs: List[torch.cuda.CUDAGraph] = []
        self.losses: List[torch.Tensor] = []
        self.launches_per_step = 0
        pool = None
        for bufs in self.buffer_sets:
            g = torch.cuda.CUDAGraph()
            n0 = L.launch_count()
            with torch.cuda.graph(g, pool=pool):
                loss = step_fn(*bufs)
            self.launches_per_step = L.launch_count() - n0   # jmt kernels recorded into the graph
            pool = g.pool()
            self.graphs.append(g)
            self.losses.append(loss)

    def replay(self, which: int = 0) -> torch.Tensor:
        """Launch the graph bound to buffer set `which`; returns its (static) device loss tensor."""
        self.graphs[which].replay()
        # the replayed optimizer step changed the parameters without touching their `_version`: eager forwards that follow
        # (validation between training epochs) must re-cast their bf16 operand copies
        from .engine import bump_param_generation
        bump_param_generation()
        return self.losses[which]
