#!/usr/bin/env python3
"""Multi-GPU check, run under torchrun on >= 2 GPUs (not collected by pytest: the driver's GPU tier has one GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py

(a) slice equivalence: the concatenation of the per-rank outputs equals the single-GPU output bit for bit;
(b) the statistics vector all-reduced over NCCL -- through torch.distributed AND through the C-ABI
    ``b200ctl_stats_allreduce`` on an ``ncclComm_t`` created by ``b200ctl_nccl_comm_init`` -- equals the
    single-GPU statistics (counts exactly, sums to 1e-12);
(c) the library's own peer-memory all-reduce (``b200ctl_stats_allreduce_peer``): same, bit-identical on all ranks,
    37 windows through the four-slot ring, and the lagged form."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from test_isaacgym_b200 import _lib, synthetic as syn  # noqa: E402
from test_isaacgym_b200.pd_control import PDController  # noqa: E402
from test_isaacgym_b200.sharding import StatsReducer, env_slice, slice_rows  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    from test_isaacgym_b200.sharding import nccl_options
    dist.init_process_group("nccl", device_id=dev, pg_options=nccl_options())
    n, d = 100_003, 12
    full = syn.pd_inputs(n, d, seed=5)
    ctl = PDController(d, full.kp, full.kd, tau_max=full.tau_max, device=dev)
    lo, hi = env_slice(n, rank, world)
    st_local = _lib.stats_buffer(dev)
    part = ctl(slice_rows(full.dof_state, d, lo, hi).to(dev), full.q_target[lo:hi].to(dev), stats=st_local)

    # single-GPU truth, computed by every rank on its own device
    st_full = _lib.stats_buffer(dev)
    whole = ctl(full.dof_state.to(dev), full.q_target.to(dev), stats=st_full)
    assert torch.equal(part, whole[lo:hi]), "slice output differs from the unsharded output"

    for backend in ("torch", "abi"):
        red = StatsReducer(backend, dev)
        st = st_local.clone()
        ev = red.all_reduce(st)
        if ev is not None:
            ev.synchronize()
        torch.cuda.synchronize(dev)
        assert st[0].item() == n and st[3].item() == st_full[3].item() and st[4].item() == st_full[4].item(), backend
        assert torch.allclose(st[1:3], st_full[1:3], rtol=1e-12), backend
        red.close()
    # (c) the library's own all-reduce over NVLink peer memory: exact counts, sums to 1e-12, BIT-IDENTICAL on all ranks;
    #     many windows back to back (ring of four slots), then the lagged form (window w yields the sum of w - 1)
    from test_isaacgym_b200.sharding import PeerStatsReducer
    red = PeerStatsReducer(dev)
    for w in range(37):
        st = st_local.clone() * (w + 1)
        red.all_reduce(st)
        torch.cuda.synchronize(dev)
        assert st[0].item() == n * (w + 1), ("peer", w, st)
        assert torch.allclose(st[1:3], st_full[1:3] * (w + 1), rtol=1e-12), ("peer", w)
        ref = st.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(ref, st), "ranks disagree on the bits of the reduced vector"
    assert red.timeouts() == 0
    red.close()
    lag = PeerStatsReducer(dev, lagged=True)
    for w in range(23):
        st = st_local.clone() * (w + 1)
        lag.all_reduce(st)
        torch.cuda.synchronize(dev)
        assert st[0].item() == n * w, ("peer-lagged", w, st)          # the previous window's sum (zeros at w = 0)
    assert lag.timeouts() == 0
    lag.close()
    dist.barrier()
    if rank == 0:
        print(f"multi-GPU check OK on {world} GPUs: slices bit-exact, stats all-reduce (torch NCCL, C-ABI NCCL, b200ctl peer-memory kernel incl. lagged form) exact")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
