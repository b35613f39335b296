"""Where do the ~28 us go that a 20-step timed region (the driver's `--steps 20 --warmup 5`) reads above 20 x the
400-step average?  Times the bench's own PD loop (same PdWorkload, same bound calls) for K in {20, 100, 400}, ten
repeats each, and once with an event between every pair of steps."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from test_isaacgym_b200 import _lib  # noqa: E402
from test_isaacgym_b200.sharding import StatsWindow  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
wl = bench.PdWorkload(dev, bench.ENVS_PER_GPU, seed=1000)
wl.bind(StatsWindow(dev, None, 16))


def loop(k):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    s.record()
    for i in range(k):
        wl.step(i)
    e.record()
    torch.cuda.synchronize(dev)
    return s.elapsed_time(e) * 1e3 / k


t_end = time.perf_counter() + 1.0
while time.perf_counter() < t_end:
    for i in range(208):
        wl.step(i)
    torch.cuda.synchronize(dev)
for k in (20, 100, 400, 20):
    print(k, "steps:", " ".join(f"{loop(k):.2f}" for _ in range(10)), "us/step")

# an event after every step of a 24-step region
evs = [torch.cuda.Event(enable_timing=True) for _ in range(25)]
torch.cuda.synchronize(dev)
evs[0].record()
for i in range(24):
    wl.step(i)
    evs[i + 1].record()
torch.cuda.synchronize(dev)
print("per-step (events between steps):", " ".join(f"{evs[i].elapsed_time(evs[i + 1]) * 1e3:.1f}" for i in range(24)))

# host cost of enqueuing
t0 = time.perf_counter()
for i in range(400):
    wl.step(i)
t1 = time.perf_counter()
torch.cuda.synchronize(dev)
print(f"host enqueue cost {1e6 * (t1 - t0) / 400:.2f} us/step")

# the same 20 steps with the zero_() of the window kept out (every = 10**9)
wl.bind(StatsWindow(dev, None, 10 ** 9))
print("no window fill, 20 steps:", " ".join(f"{loop(20):.2f}" for _ in range(10)))
# a 20-step graph (one launch)
from test_isaacgym_b200.graph import StepGraph  # noqa: E402
st = _lib.stats_buffer(dev)
calls = [wl.ctl.bind(wl.state[i % wl.sets], wl.tgt[i % wl.sets], wl.out[i % wl.sets], stats=st) for i in range(20)]
g = StepGraph(calls, dev)
for _ in range(5):
    g()
res = []
for _ in range(10):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    s.record()
    g()
    e.record()
    torch.cuda.synchronize(dev)
    res.append(s.elapsed_time(e) * 1e3 / 20)
print("20-step graph, one replay:", " ".join(f"{x:.2f}" for x in res))
