#!/usr/bin/env python3
"""Host-side cost of one control step (the time the launching thread spends per step), with and without the
statistics window's all-reduce.  Run plain (1 GPU) or under torchrun (N GPUs):

    python profiles/experiments/host_cost.py
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 profiles/experiments/host_cost.py
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import bench  # noqa: E402
from test_isaacgym_b200.sharding import StatsReducer, StatsWindow, nccl_options  # noqa: E402


def main():
    rank, local, world = bench.dist_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev, pg_options=nccl_options())
    for n in (1024, 1_048_576):
        wl = bench.PdWorkload(dev, n, seed=rank)
        for every, overlap in ((16, False), (16, True), (64, False), (10 ** 9, False)):
            wl.bind(StatsWindow(dev, StatsReducer("torch", dev) if world > 1 else None, every, overlap=overlap))
            for i in range(200):
                wl.step(i)
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            steps = 2000 if n == 1024 else 400
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            t0 = time.perf_counter()
            t_call = 0.0
            for i in range(steps):
                a = time.perf_counter()
                wl.calls[i % wl.sets][wl.window.cur]()
                t_call += time.perf_counter() - a
                wl.window.step_done()
            t_host = time.perf_counter() - t0
            ev1.record()
            torch.cuda.synchronize(dev)
            wl.window.finish()
            if rank == 0:
                print(f"world={world} n={n} all-reduce every {every if every < 10**9 else 'never'}{' (side stream)' if overlap else ''}: host {t_host / steps * 1e6:.1f} us/step "
                      f"(bound call {t_call / steps * 1e6:.1f} us), device {ev0.elapsed_time(ev1) / steps * 1e3:.1f} us/step", flush=True)
        del wl
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
