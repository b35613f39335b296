"""CUDA-graph capture of bound control-law calls.

At 16K-64K environments one control law is a 3-9 us kernel; issuing it through Python costs more than running
it.  ``BoundCall`` objects (``PDController.bind``, ``ServoStep.bind``, ``TaskStep.bind``, ``bind_control_osc`` ...)
are stream-ordered and allocation-free, so a whole control step -- e.g. ``[task, osc]`` of the Franka pick loop --
can be captured once and replayed with a single launch per step.  Kernels inside the graph keep their
programmatic-dependent-launch edges, so consecutive laws overlap their launch latency.
"""
from __future__ import annotations

import torch


class StepGraph:
    """Capture ``calls`` (zero-argument bound calls, executed in order) into one CUDA graph.

    >>> step = StepGraph([task.bind(dpose, pos_action[:, 7:9]), ctl.bind_control_osc(dpose, effort_action[:, :7])])  # doctest: +SKIP
    >>> step()          # one graph launch on the current stream                                                        # doctest: +SKIP

    Side effect of construction: a kernel's first launch (module load, shared-memory opt-in, pointer validation) cannot
    happen inside capture, so every bound call is EXECUTED ``warmup`` times for real first.  The laws are in place --
    ``servo_step`` rewrites the root state, the pick kernels advance the ``hand_restart`` latch, statistics are
    accumulated -- so building a graph with ``warmup=1`` advances the simulation state by one control step.  Pass
    ``restore=[tensors...]`` to have those tensors snapshotted before and copied back after the warm-up (statistics
    buffers, latches, in-place state), or ``warmup=0`` when the same entry points have already run in this process.
    """

    def __init__(self, calls, device: torch.device | None = None, warmup: int = 1, restore=()):
        self.calls = list(calls)
        if not self.calls:
            raise ValueError("StepGraph needs at least one bound call")
        self.device = device if device is not None else self.calls[0].device
        self.stream = torch.cuda.Stream(self.device)
        self.graph = torch.cuda.CUDAGraph()
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            saved = [(t, t.clone()) for t in restore] if warmup else []
            for _ in range(warmup):          # first launches (module load, smem opt-in) must happen outside capture
                for c in self.calls:
                    c()
            for t, snap in saved:
                t.copy_(snap)
            self.stream.synchronize()
            with torch.cuda.graph(self.graph, stream=self.stream):
                for c in self.calls:
                    c()
        torch.cuda.current_stream(self.device).wait_stream(self.stream)

    def __call__(self) -> None:
        self.graph.replay()

    def __len__(self) -> int:
        return len(self.calls)
