"""Multi-GPU plumbing: contiguous env slices per rank + all-reduce of the per-step statistics.

Environments are independent: GPU ``g`` of ``G`` owns envs
``[g*N/G, (g+1)*N/G)`` and the control step moves no bytes between GPUs.  The
only exchange is a sum all-reduce of the ``float64[8]`` statistics vector the
kernels accumulate.  Three carriers, fastest first:

* inside the control kernel itself, every step (``PDController.bind(..., publish=PeerStatsReducer(...))`` ->
  ``b200ctl_pd_torque_published``): one extra CTA of each launch publishes the previous step's vector over NVLink
  peer memory and sums the rows of the step before;
* the library's stand-alone peer-memory kernel (``PeerStatsReducer.all_reduce``) every k steps (``StatsWindow``),
  in order on the control stream or on a side stream next to the control kernels (``overlap=True``);
* NCCL (``StatsReducer``), in order on the control stream (its 640-thread CTA does not fit next to persistent grids;
  ``configure_nccl_for_control_loops`` + one reserved CTA slot is the side-stream form, measured slower).
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


def configure_nccl_for_control_loops() -> None:
    """Environment for a process whose only collectives are 64-byte statistics all-reduces issued next to persistent
    control grids (call BEFORE ``init_process_group``; values already set by the user win):

    * ``NCCL_MAX_NCHANNELS=1`` / ``NCCL_MIN_NCHANNELS=1`` -- one channel, i.e. ONE NCCL CTA per collective;
    * ``NCCL_NTHREADS=128`` -- that CTA is 128 threads (<= 96 registers each): 12,288 registers, exactly the footprint
      of one CTA of the statistics-carrying PD kernel (256 threads x 48 registers).  NCCL's default 640-thread CTA needs
      an EMPTY SM, which a chain of persistent grids never offers -- the collective then starves until the control stream
      itself blocks on it (47.5 us/step instead of 32.4 at N=2, round 1);
    * with ``b200ctl_reserve_cta_slots(device, 1)`` every persistent grid leaves one such slot free, so the all-reduce
      on its side stream runs NEXT to the control kernels instead of displacing one of their CTAs into a second wave."""
    import os
    for k, v in (("NCCL_MAX_NCHANNELS", "1"), ("NCCL_MIN_NCHANNELS", "1"), ("NCCL_NTHREADS", "128")):
        os.environ.setdefault(k, v)


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
    On GPUs the reduction is enqueued in order on the producing stream, or (``overlap=True``) on a side stream ordered
    after the producing stream by an event.
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


class PeerStatsReducer:
    """Sum all-reduce of the statistics vector over NVLink peer memory -- the library's own collective
    (``b200ctl_stats_allreduce_peer``, ``csrc/peer.cu``): one 64-thread kernel per window, in order on the control
    stream.  NCCL / ``torch.distributed`` is used once, for the plumbing (all-gather of the mailboxes' IPC handles).

    ``lagged=True``: ``all_reduce`` leaves in ``stats`` the global sum of the PREVIOUS call's vector (zeros on the first
    call) and never waits for a peer -- the form a per-step exchange needs.  Same interface as ``StatsReducer``."""

    def __init__(self, device: torch.device, lagged: bool = False, timeout_s: float = 2.0):
        if device.type != "cuda":
            raise _lib.B200CtlError(-2, "PeerStatsReducer needs CUDA devices (NVLink peer memory)")
        self.device, self.lagged, self.timeout_s = device, bool(lagged), float(timeout_s)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.window = 0
        self._side = None
        self._bound = {}
        L = _lib.lib()
        self._fn = L.b200ctl_stats_allreduce_peer
        self._mine = ctypes.c_void_p()
        handle = (ctypes.c_char * 64)()
        _lib.check(L.b200ctl_peer_mailbox_create(device.index or 0, ctypes.byref(self._mine), handle))
        mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).clone()
        if self.world > 1:
            on_gpu = dist.get_backend() == "nccl"
            src = mine.to(device) if on_gpu else mine
            gathered = [torch.empty_like(src) for _ in range(self.world)]
            dist.all_gather(gathered, src)
            handles = [bytes(t.cpu().tolist()) for t in gathered]
        else:
            handles = [bytes(mine.tolist())]
        self._boxes = (ctypes.c_void_p * self.world)()
        failure = None
        for r in range(self.world):
            if r == self.rank:
                self._boxes[r] = self._mine.value
            else:
                peer = ctypes.c_void_p()
                try:
                    _lib.check(L.b200ctl_peer_mailbox_open(device.index or 0, handles[r], ctypes.byref(peer)))
                except _lib.B200CtlError as e:      # no peer access to rank r (no NVLink / IPC closed in this container)
                    failure = failure or e
                self._boxes[r] = peer.value
        if self.world > 1:
            # every mailbox is mapped everywhere before the first publish; the same collective carries the verdict, so
            # that a rank that could not map a peer does not leave the others waiting: all ranks raise together
            ok = torch.tensor([0 if failure else 1], dtype=torch.int32, device=device if dist.get_backend() == "nccl" else "cpu")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                raise _lib.B200CtlError(-2, f"peer mailboxes could not be mapped on every rank ({failure or 'another rank failed'})")
        elif failure:
            raise failure

    def all_reduce(self, stats: torch.Tensor, overlap: bool = False, zero_after: torch.Tensor | None = None,
                   out: torch.Tensor | None = None):
        """In order on the current stream by default; returns None.
        ``zero_after``: another statistics buffer the same kernel clears (the next window's accumulator).
        ``out``: out-of-place -- the sum goes to ``out`` and ``stats`` is cleared by the kernel.  With ``overlap=True``
        (needs ``out``) the kernel runs on the reducer's side stream, ordered after everything already on the current
        stream, and the returned event marks its completion: wait for it before ``stats`` is accumulated into again
        or ``out`` is read.  The 64-thread CTA fits next to the persistent control grids (no slot to reserve)."""
        key = (stats.data_ptr(), zero_after.data_ptr() if zero_after is not None else 0, out.data_ptr() if out is not None else 0)
        args = self._bound.get(key)
        if args is None:       # marshalled once per buffer combination: the step loop alternates between two
            args = [self._boxes, self.rank, self.world, 0, 1 if self.lagged else 0, _lib.stats_arg(stats, self.device),
                    min(stats.numel(), _lib.STATS_LEN), _lib.stats_arg(zero_after, self.device), _lib.stats_arg(out, self.device),
                    self.timeout_s, self.device.index or 0, None]
            self._bound[key] = args
        args[3] = self.window
        self.window += 1
        if overlap:
            if out is None:
                raise ValueError("overlap=True needs the out-of-place form (out=...)")
            if self._side is None:
                self._side = torch.cuda.Stream(self.device, priority=-1)
            self._side.wait_stream(torch.cuda.current_stream(self.device))
            args[11] = self._side.cuda_stream
            rc = self._fn(*args)
            if rc:
                _lib.check(rc)
            ev = torch.cuda.Event()
            ev.record(self._side)
            return ev
        args[11] = torch.cuda.current_stream(self.device).cuda_stream
        rc = self._fn(*args)
        if rc:
            _lib.check(rc)
        return None

    def timeouts(self) -> int:
        n = ctypes.c_uint64()
        _lib.check(_lib.lib().b200ctl_peer_mailbox_timeouts(self.device.index or 0, self._mine, ctypes.byref(n)))
        return int(n.value)

    def wait(self) -> None:
        if self._side is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._side)

    @property
    def stream(self):
        return self._side

    def close(self) -> None:
        L = _lib.lib()
        torch.cuda.synchronize(self.device)
        if self.world > 1 and dist.is_initialized():
            dist.barrier()          # nobody unmaps a mailbox a peer may still write to
        for r in range(self.world):
            if self._boxes[r] and r != self.rank:
                L.b200ctl_peer_mailbox_close(self.device.index or 0, ctypes.c_void_p(self._boxes[r]), 1)
                self._boxes[r] = None
        if self._mine:
            L.b200ctl_peer_mailbox_close(self.device.index or 0, self._mine, 0)
            self._mine = ctypes.c_void_p()


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
        self.reduced = None           # out-of-place results of the overlapped peer form
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
        fused_zero = isinstance(self.reducer, PeerStatsReducer)
        if fused_zero and self.overlap:
            # out of place on the side stream: the kernel clears `done` after reading it and writes the sum to
            # reduced[done]; the control stream only waits for it before it accumulates into `done` again
            if self.reduced is None:
                self.reduced = [_lib.stats_buffer(self.device), _lib.stats_buffer(self.device)]
            self.events[done] = self.reducer.all_reduce(self.bufs[done], overlap=True, out=self.reduced[done])
            self.last_reduced = self.reduced[done]
            self.cur ^= 1
            ev = self.events[self.cur]
            if ev is not None:
                torch.cuda.current_stream(self.device).wait_event(ev)
                self.events[self.cur] = None
            return
        if fused_zero:                # one kernel: reduce `done` over the ranks and clear the other buffer
            self.reducer.all_reduce(self.bufs[done], zero_after=self.bufs[done ^ 1])
        elif self.reducer is not None:
            self.events[done] = self.reducer.all_reduce(self.bufs[done], overlap=self.overlap)
        self.last_reduced = self.bufs[done]
        self.cur ^= 1
        ev = self.events[self.cur]
        if ev is not None:            # the buffer we are about to reuse was reduced `every` steps ago
            torch.cuda.current_stream(self.device).wait_event(ev)
            self.events[self.cur] = None
        if not fused_zero:
            self.bufs[self.cur].zero_()
        # note: `last_reduced` must be read (after waiting on its event) before the window after next starts

    def finish(self) -> None:
        for ev in self.events:
            if ev is not None:
                torch.cuda.current_stream(self.device).wait_event(ev)
