"""CUDA-graph capture of one whole iteration of the path (render -> loss -> backward -> all-reduce -> optimizer).

A training iteration at production size is ~70 kernel launches, ~45 of them a few microseconds long (the loss,
the scalar networks, Adam); with few rays per GPU -- BASELINE.json configs[3] sharded over 8 GPUs is 512 rays each
-- the step is launch-bound (profiles/r02_scaling).  Nothing on the path synchronises with the host or branches
on device values, so the whole iteration captures into one graph and replays with a single launch.

    step = GraphedStep(train_iteration, [rays_o, rays_d, depth_scale, true_edge])
    for batch in loader:
        loss, = step(*batch)              # copies the batch into the static inputs, replays, returns static outputs

Rules (torch.cuda.graph semantics):
  * everything ``fn`` reads must be a static input, a parameter, or a constant: Python scalars (``cos_anneal_ratio``,
    ``near``/``far`` floats, learning rates held as floats) are FROZEN into the graph -- re-capture when they change
    (the reference anneals ``cos_anneal_ratio`` only during the first ``anneal_end`` iterations);
  * optimizers need ``capturable=True``; gradients must be ``None`` at capture (``zero_grad(set_to_none=True)``)
    so that the captured backward allocates them from the graph's pool;
  * the renderer must draw its stratified offsets on the device (``renderer.perturb_on_device = True``);
  * the device status word (NaN guards) is not polled inside the graph: call ``renderer.check_numerics()`` when
    convenient;  ``UDFNetwork`` modules passed as ``refold`` are invalidated after every replay, because the
    optimizer inside the graph changes the parameters behind the fold cache's back.
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Sequence, Tuple

import torch

from . import _cabi as C
from . import ops


class GraphedStep:
    def __init__(self, fn: Callable[..., Sequence[torch.Tensor]], example_inputs: Sequence[torch.Tensor],
                 warmup: int = 3, refold: Iterable = (), pool=None):
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedStep needs a CUDA device (no CPU fallback)")
        self.fn = fn
        self.refold = list(refold)
        self.inputs: List[torch.Tensor] = [t.detach().clone() for t in example_inputs]
        timed, C.timed_call = C.timed_call, None        # (bench hook: timing events cannot be captured)
        ops.graph_prepare = True                        # what a graph would freeze is refused from the warm-up on
        try:
            # warm-up on a side stream (torch.cuda.graph requirement): first-call attribute setup, lazy
            # allocations, optimizer state
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(max(warmup, 1)):
                    fn(*self.inputs)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self._invalidate()                            # the fold of the weights is part of the captured iteration
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, pool=pool):
                out = fn(*self.inputs)
        finally:
            C.timed_call = timed
            ops.graph_prepare = False
        self.outputs: Tuple[torch.Tensor, ...] = tuple(out) if isinstance(out, (tuple, list)) else (out,)
        self._invalidate()

    def _invalidate(self) -> None:
        for m in self.refold:
            m.invalidate()

    def __call__(self, *inputs: torch.Tensor) -> Tuple[torch.Tensor, ...]:
        if len(inputs) != len(self.inputs):
            raise ValueError(f"GraphedStep: {len(self.inputs)} inputs captured, {len(inputs)} given")
        for dst, src in zip(self.inputs, inputs):
            if src is not dst:
                if src.shape != dst.shape or src.dtype != dst.dtype:
                    raise ValueError("GraphedStep: input shape/dtype differs from the captured one")
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        self._invalidate()
        return self.outputs
