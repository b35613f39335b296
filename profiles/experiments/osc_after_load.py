#!/usr/bin/env python3
"""Why does the 262,144-env OSC entry read ~7 % slower inside bench.py than in quick_time.py?  Times it (a) on a fresh
process, (b) after 3 s of the headline PD loop, (c) after a further 5 s idle, sampling the SM clock each time."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
import bench  # noqa: E402
from test_isaacgym_b200 import synthetic as syn  # noqa: E402
import test_isaacgym_b200.franka_cube_ik_osc as ctl  # noqa: E402
from test_isaacgym_b200.sharding import StatsWindow  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    n = 262_144
    fi = syn.franka_inputs(n, seed=3)
    calls, keep = [], []
    for _ in range(2):
        d = fi.__class__(**{k: (v.to(dev).clone() if isinstance(v, torch.Tensor) else v) for k, v in fi.__dict__.items()})
        o = torch.zeros(n, 9, device=dev)
        ctl.bind(j_eef=d.j_eef, mm=d.mm, dof_pos=d.dof_pos, dof_vel=d.dof_vel, default_dof_pos_tensor=d.default_dof_pos,
                 num_envs=n, precision=0)
        ctl.bind_hand(d.rb_states, d.hand_idxs)
        calls.append(ctl.bind_control_osc(d.dpose, o[:, :7]))
        keep.append((d, o))

    def timed(tag):
        with bench.ClockSampler(0) as cs:
            ts = [bench.graph_time(calls, dev, 40, runs=1) * 1e3 for _ in range(6)]
        print(f"{tag}: {', '.join('%.2f' % t for t in ts)} us; clocks {cs.summary()}", flush=True)

    timed("fresh process")
    wl = bench.PdWorkload(dev, 1_048_576, seed=1)
    wl.bind(StatsWindow(dev, None, 16))
    t_end = time.perf_counter() + 3.0
    i = 0
    while time.perf_counter() < t_end:
        for _ in range(200):
            wl.step(i)
            i += 1
        torch.cuda.synchronize(dev)
    timed("right after 3 s of the PD loop")
    time.sleep(5.0)
    timed("after a further 5 s idle")
    del wl
    torch.cuda.empty_cache()
    timed("PD buffers freed")


if __name__ == "__main__":
    main()
