import sys, os
sys.path.insert(0, "/root/repo")
import torch, bench
from test_isaacgym_b200 import synthetic as syn
from test_isaacgym_b200.pd_control import PDController
dev = torch.device("cuda", 0)
for d in (12, 9, 7, 8):
    n = 1_048_576
    pi = syn.pd_inputs(n, d, seed=1)
    c = PDController(d, pi.kp, pi.kd, tau_max=pi.tau_max, device=dev)
    sets = 4
    st = [pi.dof_state.to(dev).clone() for _ in range(sets)]
    tg = [pi.q_target.to(dev).clone() for _ in range(sets)]
    ou = [torch.empty(n, d, device=dev) for _ in range(sets)]
    t = bench.graph_time([c.bind(st[k], tg[k], ou[k]) for k in range(sets)], dev, 20) * 1e3
    print(f"D={d}: {t:.2f} us  -> {n*d*16/t/1e3:.0f} GB/s", flush=True)
