#!/usr/bin/env python3
"""Multi-GPU check, run under torchrun on >= 2 GPUs (tests/test_gpu_multi.py launches it on two GPUs when the box has them):

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
    # (d) statistics exchanged inside the PD kernel (b200ctl_pd_torque_published, publisher CTA): two alternating
    #     accumulators; after step s, `reduced` holds the global sum of step s - 2; also through a CUDA graph (the window
    #     counter lives in the mailbox, so replays stay in step)
    from test_isaacgym_b200.graph import StepGraph
    pub = PeerStatsReducer(dev, lagged=True)
    acc, reduced = [_lib.stats_buffer(dev), _lib.stats_buffer(dev)], _lib.stats_buffer(dev)
    st_in, tg_in = slice_rows(full.dof_state, d, lo, hi).to(dev), full.q_target[lo:hi].to(dev)
    n_loc = hi - lo                                # D = 12: N * D % 4 == 0, the vector path the published form needs
    out_p = torch.empty(n_loc, d, device=dev)
    calls = [ctl.bind(st_in, tg_in, out_p, stats=acc[k], stats_prev=acc[k ^ 1], publish=pub, reduced=reduced) for k in (0, 1)]
    n_glob = torch.tensor([n_loc], device=dev, dtype=torch.float64)
    dist.all_reduce(n_glob)
    for s_ in range(9):
        calls[s_ & 1]()
        torch.cuda.synchronize(dev)
        assert acc[s_ & 1][0].item() == n_loc, "this step's accumulator holds this step"
        assert acc[(s_ & 1) ^ 1].abs().sum().item() == 0.0, "the publisher CTA must clear the previous step's accumulator"
        # step s publishes step s - 1 and yields the global sum of step s - 2
        assert reduced[0].item() == (0 if s_ < 2 else n_glob.item()), ("published", s_, reduced)
    assert torch.equal(out_p, whole[lo:hi])
    g = StepGraph([calls[1], calls[0]], dev, warmup=0)          # step 9 is odd: the alternation continues inside the graph
    for _ in range(5):
        g()
    torch.cuda.synchronize(dev)
    assert reduced[0].item() == n_glob.item() and pub.timeouts() == 0
    assert torch.allclose(reduced[1:3], st_full[1:3], rtol=1e-12) and reduced[3].item() == st_full[3].item()
    ref = reduced.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(ref, reduced), "ranks disagree on the published sum"
    pub.close()
    dist.barrier()
    if rank == 0:
        print(f"multi-GPU check OK on {world} GPUs: slices bit-exact, stats all-reduce (torch NCCL, C-ABI NCCL, b200ctl peer-memory kernel incl. lagged form, statistics published by the PD kernel) exact")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
