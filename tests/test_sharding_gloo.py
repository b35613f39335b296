"""world_size-2 gloo test of the multi-GPU host logic (env slices + statistics all-reduce) on CPU.

The kernels cannot run here; each rank fills its statistics vector with the oracle on its own env
slice, and the test checks (a) slices tile the env range, (b) the all-reduced vector equals the
single-process vector, (c) concatenated slice outputs equal the unsharded output bit for bit."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from test_isaacgym_b200.sharding import env_slice, slice_rows, rebase_index, StatsReducer
from test_isaacgym_b200 import synthetic as syn
from oracle import pd as opd


def test_env_slice_tiles_the_range():
    for n in (0, 1, 7, 8, 1024, 1_048_576, 1_000_003):
        for w in (1, 2, 3, 4, 8):
            edges = [env_slice(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        env_slice(8, 2, 2)


def test_rebase_index():
    idx = torch.arange(8) * 13 + 10
    s, e = env_slice(8, 1, 2)
    assert torch.equal(rebase_index(idx[s:e], 13, s), torch.arange(4) * 13 + 10)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = syn.pd_inputs(n, 12, seed=21)
        s, e = env_slice(n, rank, world)
        ds = slice_rows(full.dof_state, 12, s, e)
        tau = opd.pd_torque(ds, full.q_target[s:e], full.kp, full.kd, tau_max=full.tau_max)
        stats = opd.pd_stats(tau, full.tau_max)
        red = StatsReducer(backend="torch", device=None)
        red.all_reduce(stats)
        red.wait()
        ret[rank] = (tau, stats)
    finally:
        dist.destroy_process_group()


def test_world2_gloo_slices_and_stats():
    n, world = 1001, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, ret), nprocs=world, join=True)
    full = syn.pd_inputs(n, 12, seed=21)
    tau = opd.pd_torque(full.dof_state, full.q_target, full.kp, full.kd, tau_max=full.tau_max)
    ref_stats = opd.pd_stats(tau, full.tau_max)
    cat = torch.cat([ret[r][0] for r in range(world)])
    assert torch.equal(cat, tau)                                  # slice-equivalence, bit-exact
    for r in range(world):
        assert torch.allclose(ret[r][1], ref_stats, rtol=1e-12)   # every rank holds the global stats
    assert ret[0][1][0] == n
