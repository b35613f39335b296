#!/usr/bin/env python3
"""Why the sustained PD loop reads 0.93-0.97 of the copy bandwidth: sustained runs (2 s each) of
   copy        torch b.copy_(a) over 1 Gi bf16 elements (the MEASURED_PEAKS.json measurement)
   pd          PD law, 1,048,576 envs x 12, no statistics
   pd+stats    the headline instantiation (saturation + statistics epilogue)
   pd+all      every flag (wrap, clamp, velocity targets) with statistics
as 20-step CUDA graphs over 4 rotating buffer sets, with the SM clock, power draw and throttle reasons sampled through NVML
DURING each run, and the achieved GB/s over the run's last second.

    python profiles/experiments/pd_power_probe.py
"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import pynvml  # noqa: E402
import bench  # noqa: E402
from test_isaacgym_b200 import _lib as L, synthetic as syn  # noqa: E402
from test_isaacgym_b200.graph import StepGraph  # noqa: E402
from test_isaacgym_b200.pd_control import PDController  # noqa: E402


class Sampler:
    def __init__(self, index):
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        self.rows, self.stop = [], False

    def __enter__(self):
        self.rows, self.stop = [], False
        self.t = threading.Thread(target=self.loop, daemon=True)
        self.t.start()
        return self

    def loop(self):
        while not self.stop:
            self.rows.append((pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM),
                              pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_MEM),
                              pynvml.nvmlDeviceGetPowerUsage(self.h) / 1e3,
                              pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)))
            time.sleep(0.02)

    def __exit__(self, *exc):
        self.stop = True
        self.t.join()

    def summary(self):
        rows = self.rows[len(self.rows) // 2:]          # second half of the run: the settled state
        med = lambda xs: sorted(xs)[len(xs) // 2]
        reasons = 0
        for r in rows:
            reasons |= r[3]
        return f"sm {med([r[0] for r in rows])} MHz, mem {med([r[1] for r in rows])} MHz, {med([r[2] for r in rows]):.0f} W, reasons 0x{reasons:x}"


def sustained(fn, bytes_per_call, device, seconds=2.0):
    ev = lambda: torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds - 1.0:
        fn()
        torch.cuda.synchronize(device)
    s, e, reps = ev(), ev(), 0
    t0 = time.perf_counter()
    s.record()
    while time.perf_counter() - t0 < 1.0:
        for _ in range(8):
            fn()
        reps += 8
        torch.cuda.synchronize(device) if reps % 64 == 0 else None
    e.record()
    torch.cuda.synchronize(device)
    ms = s.elapsed_time(e) / reps
    return ms, bytes_per_call / ms / 1e6


def main():
    libs = [a for a in sys.argv[1:] if a.endswith(".so")]
    quick = "--quick" in sys.argv or bool(libs)
    if libs:
        L.LIB_PATH, L._lib = os.path.abspath(libs[0]), None
        print("library:", L.LIB_PATH, flush=True)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    n, d, sets, block = bench.ENVS_PER_GPU, bench.NUM_DOFS, 4, 20
    pi = syn.pd_inputs(n, d, seed=1000, gain_set="B")
    st = [pi.dof_state.to(dev)] + [pi.dof_state.to(dev).roll(7 * (k + 1), 0).contiguous() for k in range(sets - 1)]
    tg = [pi.q_target.to(dev)] + [pi.q_target.to(dev).roll(3 * (k + 1), 0).contiguous() for k in range(sets - 1)]
    qd = [torch.randn(n, d, device=dev) for _ in range(sets)]
    out = [torch.empty(n, d, device=dev) for _ in range(sets)]
    sb = L.stats_buffer(dev)
    plain = PDController(d, pi.kp, pi.kd, tau_max=pi.tau_max, device=dev)
    lo, hi = torch.full((d,), -3.0), torch.full((d,), 3.0)
    full = PDController(d, pi.kp, pi.kd, tau_max=pi.tau_max, q_lo=lo, q_hi=hi, wrap_angle=True, clamp_target=True, device=dev)
    a = torch.zeros(1 << 30, dtype=torch.bfloat16, device=dev)
    b = torch.empty_like(a)
    variants = [("copy", lambda: b.copy_(a), 4.0 * (1 << 30))]
    if "--copy-data" in sys.argv:       # is the copy's power a property of the DATA?  zeros vs random bits vs the PD inputs' floats
        ar = torch.randint(-32768, 32767, (1 << 30,), dtype=torch.int16, device=dev).view(torch.bfloat16)
        af = torch.randn(1 << 29, device=dev).view(torch.bfloat16)          # fp32 N(0,1) values, viewed as pairs of bf16
        variants += [("copy_rand", lambda: b.copy_(ar), 4.0 * (1 << 30)), ("copy_f32n", lambda: b.copy_(af), 4.0 * (1 << 30))]
        for name, fn, nbytes in variants:
            with Sampler(0) as smp0:
                ms, gbs = sustained(fn, nbytes, dev)
            print(f"{name:9s} {ms * 1e3:8.2f} us per copy  {gbs:7.1f} GB/s   {smp0.summary()}", flush=True)
        return

    def graph_of(ctl, **kw):
        calls = [ctl.bind(st[k % sets], tg[k % sets], out[k % sets], **{key: (v[k % sets] if isinstance(v, list) else v) for key, v in kw.items()})
                 for k in range(block)]
        for c in calls[:sets]:
            c()
        return StepGraph(calls, dev, warmup=0)
    variants.append(("pd", graph_of(plain), block * n * 192.0))
    variants.append(("pd+stats", graph_of(plain, stats=sb), block * n * 192.0))
    if quick:
        variants = variants[1:]
    else:
        variants.append(("pd+all", graph_of(full, qd_target=qd, stats=sb), block * n * 240.0))
    smp = Sampler(0)
    for rnd in range(1 if quick else 2):
        for name, fn, nbytes in variants:
            with smp:
                ms, gbs = sustained(fn, nbytes, dev)
            per = ms * 1e3 / (1 if name.startswith("copy") else block)
            print(f"round {rnd} {name:9s} {per:8.2f} us per {'copy' if name == 'copy' else 'step'}  {gbs:7.1f} GB/s   {smp.summary()}", flush=True)
            time.sleep(0.5)


if __name__ == "__main__":
    main()
