"""Multi-GPU plumbing: contiguous env slices per rank + all-reduce of the per-step statistics.

Environments are independent: GPU ``g`` of ``G`` owns envs
``[g*N/G, (g+1)*N/G)`` and the control step moves no bytes between GPUs.  The
only collective is a sum all-reduce of the ``float64[8]`` statistics vector the
kernels accumulate -- issued on a side stream so it never gates the next step.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _lib


def env_slice(num_envs: int, rank: int, world_size: int) -> tuple[int, int]:
    """[start, stop) of the contiguous env slice owned by ``rank`` (remainder spread over the first ranks)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, rem = divmod(num_envs, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def slice_rows(t: torch.Tensor, rows_per_env: int, start: int, stop: int) -> torch.Tensor:
    """Rows of a flat Isaac Gym tensor ((N*rows_per_env, C)) that belong to envs [start, stop)."""
    return t[start * rows_per_env: stop * rows_per_env]


def rebase_index(index: torch.Tensor, rows_per_env: int, start: int) -> torch.Tensor:
    """Global rigid-body indices (e.g. ``hand_idxs``) rebased to a slice starting at env ``start``."""
    return index - start * rows_per_env


def nccl_options():
    """``pg_options`` for ``dist.init_process_group("nccl", ...)`` in a process that runs b200ctl step loops: the
    collectives go to a high-priority CUDA stream.  The control kernels are persistent grids chained by programmatic
    dependent launch -- the next step's CTAs take every SM slot the moment the previous step's CTAs retire -- so a
    normal-priority all-reduce kernel does not get a slot until the control stream blocks on the statistics window's
    event, and then still waits for the peer ranks to reach the same point.  (A non-torch host does the same by
    calling ``b200ctl_stats_allreduce`` on a stream created with ``cudaStreamCreateWithPriority``.)"""
    from torch.distributed import ProcessGroupNCCL
    opts = ProcessGroupNCCL.Options()
    opts.is_high_priority_stream = True
    return opts


class StatsReducer:
    """Sum all-reduce of statistics vectors across ranks.

    backend "torch": ``torch.distributed.all_reduce`` on whatever process group is initialised (NCCL on GPUs,
    gloo in the CPU tests).  backend "abi": ``b200ctl_stats_allreduce`` on an ``ncclComm_t`` created through
    the C ABI (unique id broadcast over the torch process group) -- the path a non-torch host would use.
    On GPUs the reduction runs on a side stream ordered after the producing stream by an event.
    """

    def __init__(self, backend: str = "torch", device: torch.device | None = None):
        self.backend = backend
        self.device = device
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self._comm = ctypes.c_void_p()
        self._side = torch.cuda.Stream(device, priority=-1) if (device is not None and device.type == "cuda") else None
        if backend == "abi" and self.world > 1:
            L = _lib.lib()
            uid = torch.zeros(128, dtype=torch.uint8)
            if self.rank == 0:
                buf = (ctypes.c_char * 128)()
                _lib.check(L.b200ctl_nccl_unique_id(buf))
                uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
            uid = uid.to(device) if dist.get_backend() == "nccl" else uid
            dist.broadcast(uid, src=0)
            raw = bytes(uid.cpu().tolist())
            with torch.cuda.device(device):
                _lib.check(L.b200ctl_nccl_comm_init(ctypes.byref(self._comm), self.world, raw, self.rank))

    def all_reduce(self, stats: torch.Tensor, overlap: bool = False):
        """In-place sum over ranks.

        ``overlap=False`` (default): the reduction is enqueued IN ORDER on the current stream, between two control
        steps; returns None.  ``overlap=True``: it is enqueued on the side stream (ordered after everything already on
        the current stream) and a ``torch.cuda.Event`` marking its completion is returned: wait on it
        (``current_stream().wait_event(ev)`` / ``ev.synchronize()``) before reading or re-zeroing ``stats``; kernels
        of later steps must accumulate into a DIFFERENT buffer meanwhile (see ``StatsWindow``).

        In order is the default because the control kernels are ONE-WAVE persistent grids: a collective kernel that
        is resident next to them (waiting for its peers, which may lag by up to a window) displaces a few of their CTAs
        into a second wave and every step it overlaps takes ~1.6x as long -- measured at N=2, k=16: 46.5 us/step
        overlapped vs 31.6 us/step without any collective (``profiles/experiments/host_cost.py``).  In order, the
        ranks re-synchronise every window and the cost is the collective's own latency once per k steps."""
        if self.world == 1:
            return None
        if overlap and self._side is not None:
            self._side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self._side):
                self._reduce(stats)
                ev = torch.cuda.Event()
                ev.record(self._side)
            stats.record_stream(self._side)
            return ev
        self._reduce(stats)
        return None

    def _reduce(self, stats):
        if self.backend == "abi":
            _lib.check(_lib.lib().b200ctl_stats_allreduce(self._comm, ctypes.c_void_p(stats.data_ptr()), stats.numel(),
                                                          _lib.stream_ptr(self.device)))
        else:
            dist.all_reduce(stats, op=dist.ReduceOp.SUM)

    def wait(self) -> None:
        if self._side is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._side)

    @property
    def stream(self):
        return self._side

    def close(self) -> None:
        if self._comm:
            _lib.lib().b200ctl_nccl_comm_destroy(self._comm)
            self._comm = ctypes.c_void_p()


class StatsWindow:
    """Double-buffered statistics accumulator: kernels add into the current buffer; every ``every`` steps the
    finished buffer is all-reduced -- in order on the control stream by default, or on the reducer's side stream with
    ``overlap=True`` while the next steps accumulate into the other buffer (see ``StatsReducer.all_reduce`` for why
    overlapping is not the default) -- and stays readable as ``last_reduced`` for a whole window."""

    def __init__(self, device: torch.device, reducer: StatsReducer | None, every: int = 16, overlap: bool = False):
        self.bufs = [_lib.stats_buffer(device), _lib.stats_buffer(device)]
        self.events = [None, None]
        self.cur = 0
        self.reducer, self.every, self.device, self.overlap = reducer, max(1, every), device, overlap
        self.last_reduced = None      # most recent globally-reduced window (device tensor)
        self._steps = 0

    @property
    def current(self) -> torch.Tensor:
        return self.bufs[self.cur]

    def step_done(self) -> None:
        """Call once per control step, after the kernel that accumulated into ``current`` was launched."""
        self._steps += 1
        if self._steps % self.every:
            return
        done = self.cur
        if self.reducer is not None:
            self.events[done] = self.reducer.all_reduce(self.bufs[done], overlap=self.overlap)
        self.last_reduced = self.bufs[done]
        self.cur ^= 1
        ev = self.events[self.cur]
        if ev is not None:            # the buffer we are about to reuse was reduced `every` steps ago
            torch.cuda.current_stream(self.device).wait_event(ev)
            self.events[self.cur] = None
        self.bufs[self.cur].zero_()
        # note: `last_reduced` must be read (after waiting on its event) before the window after next starts

    def finish(self) -> None:
        for ev in self.events:
            if ev is not None:
                torch.cuda.current_stream(self.device).wait_event(ev)
