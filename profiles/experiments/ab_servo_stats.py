#!/usr/bin/env python3
"""A/B of the statistics-carrying servo step over several builds of the library in ONE process (one CUDA context, one
set of buffers; interleaved rounds so clock drift hits every variant alike).

    python profiles/experiments/ab_servo_stats.py [variant.so ...]      # the in-tree build is always the first entry
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from test_isaacgym_b200 import _lib as L, synthetic as syn  # noqa: E402
from test_isaacgym_b200.servo_step import ServoStep, PRECISION_FAST  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    n = 1_048_576
    base = syn.servo_root_state(n, seed=2).to(dev)
    bufs = [base.clone() for _ in range(3)]
    sb = L.stats_buffer(dev)
    libs = [("in-tree", L.LIB_PATH)] + [(os.path.basename(p), os.path.abspath(p)) for p in sys.argv[1:]]
    calls = {}
    want = {}
    for name, path in libs:
        L.LIB_PATH, L._lib = path, None          # bind this variant's entry point
        for tag, prec in (("ref", 0), ("fast", PRECISION_FAST)):
            step = ServoStep(1600, 900, precision=prec)
            for stag in ("", "+stats"):
                # same bits as the in-tree build (state and statistics vector); ragged size: bulk and loop-staged tiles
                chk, st = base[: n - 37].clone(), L.stats_buffer(dev)
                step(chk, stats=st if stag else None)
                if name == "in-tree":
                    want[tag + stag] = (chk, st)
                else:
                    w = want[tag + stag]
                    same = torch.equal(chk, w[0]) and torch.allclose(st, w[1], rtol=1e-12, atol=0)
                    print(f"{name:28s} servo_{tag + stag:10s} == in-tree: {same}", flush=True)
                calls[name, tag + stag] = [step.bind(b, **({"stats": sb} if stag else {})) for b in bufs]
    best = {k: 1e9 for k in calls}
    for _ in range(2):
        for k, c in calls.items():
            best[k] = min(best[k], bench.graph_time(c, dev, 20, runs=2, warm_ms=10.0) * 1e3)
    for (name, tag), t in best.items():
        print(f"{name:28s} servo_{tag:10s} n={n}: {t:.2f} us", flush=True)


if __name__ == "__main__":
    main()
