#!/usr/bin/env python3
"""OSC / fused pick step: one thread per env vs an adjacent lane pair per env, over launch sizes.

    python profiles/lanes_sweep.py [--sizes 4096,16384,65536,262144]

CUDA-graph replays of bound calls over rotating buffer sets larger than L2 (bench.py's `graph_time`).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from bench import graph_time  # noqa: E402
from test_isaacgym_b200 import _lib, synthetic as syn  # noqa: E402
import test_isaacgym_b200.franka_cube_ik_osc as ctl  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="4096,16384,32768,65536,131072,262144")
    ap.add_argument("--pick", action="store_true", help="also time the fused pick step")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    res = {}
    for n in [int(x) for x in a.sizes.split(",")]:
        sets = max(2, min(8, (200 << 20) // (n * 2900)))        # ~2.9 KB of gym tensors per env
        fi, ti = syn.franka_inputs(n, seed=3), syn.franka_task_inputs(n, seed=4)
        keep = []
        for _ in range(sets):
            d = fi.__class__(**{k: (v.to(dev).clone() if isinstance(v, torch.Tensor) else v) for k, v in fi.__dict__.items()})
            t = ti.__class__(**{k: (v.to(dev).clone() if isinstance(v, torch.Tensor) else v) for k, v in ti.__dict__.items()})
            keep.append((d, t, torch.zeros(n, 9, device=dev), torch.zeros(n, 9, device=dev)))
        for chain, ctag in ((0, "fp64"), (1, "fp32")):
            for lanes, ltag in ((_lib.LANES_ONE, "one"), (_lib.LANES_PAIR, "pair")):
                calls = []
                for d, t, o, pos in keep:
                    ctl.bind(j_eef=d.j_eef, mm=d.mm, dof_pos=d.dof_pos, dof_vel=d.dof_vel,
                             default_dof_pos_tensor=d.default_dof_pos, num_envs=n, precision=chain | lanes)
                    ctl.bind_hand(d.rb_states, d.hand_idxs)
                    calls.append(ctl.bind_control_osc(d.dpose, o[:, :7]))
                res[f"osc_{ctag}_{ltag}_{n}"] = round(graph_time(calls, dev, 20) * 1e3, 3)
                if a.pick:
                    calls = []
                    for d, t, o, pos in keep:
                        ctl.bind(j_eef=d.j_eef, mm=d.mm, dof_pos=t.dof_pos, dof_vel=t.dof_state[:, 1].view(n, 9, 1),
                                 default_dof_pos_tensor=d.default_dof_pos, num_envs=n, precision=chain | lanes)
                        ctl.bind_hand(t.rb_states, t.hand_idxs)
                        task = ctl.TaskStep(t.rb_states, t.box_idxs, t.hand_idxs, t.dof_pos, t.init_pos, t.init_rot,
                                            t.hand_restart, "osc")
                        calls.append(ctl.bind_pick_osc(task, o[:, :7], pos[:, 7:9]))
                    res[f"pick_{ctag}_{ltag}_{n}"] = round(graph_time(calls, dev, 20) * 1e3, 3)
        del keep
        torch.cuda.empty_cache()
    ctl.bind(precision=0)
    for k, v in res.items():
        print(f"{k:32s} {v:9.3f} us")
    print(json.dumps(res))


if __name__ == "__main__":
    main()
