#!/usr/bin/env python3
"""Benchmark of the per-environment control-law hot path (BASELINE.json metric:
"controller env-steps/sec at 64K-1M envs, 1/2/4/8 B200; % HBM / FP32 roofline").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Headline workload (config.workload): the joint PD / servo torque law on 1,048,576 envs x 12 DOF
PER GPU (BASELINE.json configs[3] at one shard per GPU; it fits one GPU, and it is the size the
north-star target is quoted on).  One "step" = one evaluation of the law over the whole batch:
one kernel launch with the fused statistics epilogue.  Envs shard as contiguous slices, one
process per GPU, no data-path collective ("scaling": "weak"); at N > 1 the float64[8] statistics
vector is all-reduced over NCCL every --stats-every steps on a side stream.

Printed keys beyond the base contract:
  roofline      dominant kernel vs measured HBM peak (MEASURED_PEAKS.json), algorithmic bytes
  cpu_baseline  the oracle port (torch-CPU restatement of the reference expression) on the host cores
  e2e           same metric through the public host-tensor API (pinned host buffers, H2D + D2H timed)
  families      per-GPU numbers for the other laws / sizes of BASELINE.json configs (rank 0, N=1)
  clocks        SM clock / throttle reasons sampled (NVML) while the kernels run
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "controller env-steps/sec"
UNIT = "env-steps/s"
ENVS_PER_GPU = 1_048_576
NUM_DOFS = 12
PD_BYTES_PER_ENV = NUM_DOFS * 16          # 8 B state + 4 B target + 4 B output per DOF
SERVO_BYTES_PER_ENV = 96
IK_BYTES_PER_ENV, IK_FLOPS = 248, 480
OSC_BYTES_PER_ENV, OSC_FLOPS = 496, 1850
FALLBACK_HBM_GBS = 6650.0                 # /opt/skills/guides/B200_PROFILING.md fallback
# measured DRAM traffic per env (ncu --set full, read + write), by family-entry prefix
DRAM_BYTES_PER_ENV = {
    "servo_step": (208, "profiles/r01_full_servo_v4.txt: 109.1 MB read per 1,048,576 envs; writes are the same rows"),
    "osc_": (970, "profiles/r01_full_osc_v4.txt: 244.4 MB read + 10 MB written per 262,144 envs"),
    "ik_": (460, "profiles/r01_full_ik_v4.txt: 115.2 MB read + 6 MB written per 262,144 envs"),
    "franka_task": (438, "profiles/r01_full_franka_task.txt: 110.4 MB read + 4.4 MB written per 262,144 envs"),
}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def workload_name():
    return f"pd_torque: {ENVS_PER_GPU} envs x {NUM_DOFS} DOF per GPU (BASELINE configs[3], one env slice per GPU)"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the benchmark runs."""

    def __init__(self, index: int, period_s: float = 0.02):
        self.index, self.period = index, period_s
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    _NAMES = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
              0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
              0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def _loop(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)
                util = nv.nvmlDeviceGetUtilizationRates(self._h).gpu
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                self.samples.append((mhz, util))
                for bit, name in self._NAMES.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=1.0)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        busy = [m for m, u in self.samples if u > 0] or [m for m, _ in self.samples]
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ helpers
def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def time_launches(fn, steps: int, warmup: int, device) -> float:
    """Device time (ms) of `steps` back-to-back calls, CUDA events on the launching stream."""
    for _ in range(warmup):
        fn()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(device)
    start.record()
    for i in range(steps):
        fn(i)
    end.record()
    torch.cuda.synchronize(device)
    return start.elapsed_time(end)


class PdWorkload:
    """Rotating buffer sets (inputs + outputs of one set = 201 MB; the rotation keeps every step's data out of
    the 126 MB L2, on top of each set's inputs alone exceeding it)."""

    def __init__(self, device, num_envs, seed, sets=4, pinned_host=False):
        from test_isaacgym_b200 import synthetic as syn
        from test_isaacgym_b200.pd_control import PDController
        self.device, self.n = device, num_envs
        pi = syn.pd_inputs(num_envs, NUM_DOFS, seed=seed, gain_set="B")
        self.host = pi
        self.ctl = PDController(NUM_DOFS, pi.kp, pi.kd, tau_max=pi.tau_max, device=device)
        base_state, base_tgt = pi.dof_state.to(device), pi.q_target.to(device)
        self.state = [base_state] + [base_state.roll(7 * (k + 1), 0).contiguous() for k in range(sets - 1)]
        self.tgt = [base_tgt] + [base_tgt.roll(3 * (k + 1), 0).contiguous() for k in range(sets - 1)]
        self.out = [torch.empty(num_envs, NUM_DOFS, device=device) for _ in range(sets)]
        self.sets = sets
        self.window = None      # StatsWindow, set by the caller
        self.calls = None

    def bind(self, window):
        """Marshal one call per (buffer set, statistics buffer): the per-step host cost is one foreign call."""
        self.window = window
        self.calls = [[self.ctl.bind(self.state[k], self.tgt[k], self.out[k], stats=b) for b in window.bufs]
                      for k in range(self.sets)]

    def step(self, i=0):
        self.calls[i % self.sets][self.window.cur]()
        self.window.step_done()


# ------------------------------------------------------------------------------------------ families (rank 0, N=1)
def graph_time(calls, device, reps, warm=3, runs=3, warm_ms=30.0, stat="min"):
    """Capture `calls` (bound C-ABI calls, one kernel each) into one CUDA graph (`StepGraph`) and time `reps`
    replays with CUDA events on the replaying stream, `runs` times.  Returns ms per kernel launch: the best run
    (`stat="min"`) or the median.  The graph removes the host launch cost, which at 64K envs is larger than the
    kernels themselves.  Replays run for `warm_ms` before the first timed run: these entries follow seconds of host-side
    input generation, and an idle GPU needs tens of milliseconds to come back to its boost clock (a single timed run
    after three warm replays read 5-8 % high on the latency-bound kernels)."""
    from test_isaacgym_b200.graph import StepGraph
    g = StepGraph(calls, device)
    stream = torch.cuda.current_stream(device)
    for _ in range(warm):
        g()
    stream.synchronize()
    t_end = time.perf_counter() + warm_ms * 1e-3
    while time.perf_counter() < t_end:
        g()
        stream.synchronize()
    times = []
    for _ in range(runs):
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stream.synchronize()
        start.record(stream)
        for _ in range(reps):
            g()
        end.record(stream)
        stream.synchronize()
        times.append(start.elapsed_time(end) / (reps * len(calls)))
    times.sort()
    return times[0] if stat == "min" else times[len(times) // 2]


def family_numbers(device, peak_gbs):
    """Per-GPU throughput of the other laws / sizes named in BASELINE.json configs (not bench lines of their own).
    Every entry rotates enough buffer sets to exceed the 126 MB L2 and replays a CUDA graph of bound calls."""
    from test_isaacgym_b200 import synthetic as syn
    from test_isaacgym_b200.pd_control import PDController
    from test_isaacgym_b200.servo_step import ServoStep, PRECISION_FAST
    import test_isaacgym_b200.franka_cube_ik_osc as ctl
    out = {"_timing_note": "each entry: CUDA-graph replays of bound calls over rotating buffer sets larger than L2, "
                           "30 ms of warm replays, best of 3 timed runs of `reps` replays"}

    def record(name, n, bytes_per_env, calls, reps, flops=None):
        ms = graph_time(calls, device, reps)
        rate = n / (ms * 1e-3)
        e = {"envs": n, "us_per_step": round(ms * 1e3, 3), "env_steps_per_s": rate,
             "hbm_frac": rate * bytes_per_env / (peak_gbs * 1e9), "buffer_sets": len(calls)}
        if flops:
            e["tflops_canonical"] = rate * flops / 1e12
        # DRAM bytes per env actually moved (ncu dram__bytes_read + write at the throughput size, profiles/): the gym
        # layouts (13-float rows, a 6x7 slot of a 10x6x9 jacobian, a 7x7 corner of a 9x9 matrix) are fetched at
        # 64-byte granularity, so this exceeds the algorithmic bytes without any re-read
        for key, (b, src) in DRAM_BYTES_PER_ENV.items():
            if name.startswith(key):
                e["dram_bytes_per_env_ncu"] = b
                e["dram_frac"] = rate * b / (peak_gbs * 1e9)
                e["dram_bytes_source"] = src
        out[name] = e

    # P at C2 (65,536 x 12 = 12.6 MB per set): 24 rotating sets (302 MB)
    n, sets = 65_536, 24
    pi = syn.pd_inputs(n, NUM_DOFS, seed=1)
    c = PDController(NUM_DOFS, pi.kp, pi.kd, tau_max=pi.tau_max, device=device)
    st = [pi.dof_state.to(device).clone() for _ in range(sets)]
    tg = [pi.q_target.to(device).clone() for _ in range(sets)]
    ou = [torch.empty(n, NUM_DOFS, device=device) for _ in range(sets)]
    record("pd_65536x12", n, PD_BYTES_PER_ENV, [c.bind(st[k], tg[k], ou[k]) for k in range(sets)], reps=20)
    del st, tg, ou

    # P at C1 (1,024 x 12, BASELINE configs[0]: 197 KB per set, L2-resident whatever the rotation): one CTA wave of
    # 12 CTAs, so this is the floor of ONE graph-replayed, programmatically-serialised launch on this box -- the
    # yardstick for every C2 / C3 entry below
    n, sets = 1_024, 24
    pi = syn.pd_inputs(n, NUM_DOFS, seed=1)
    c1 = PDController(NUM_DOFS, pi.kp, pi.kd, tau_max=pi.tau_max, device=device)
    st = [pi.dof_state.to(device).clone() for _ in range(sets)]
    tg = [pi.q_target.to(device).clone() for _ in range(sets)]
    ou = [torch.empty(n, NUM_DOFS, device=device) for _ in range(sets)]
    record("pd_1024x12", n, PD_BYTES_PER_ENV, [c1.bind(st[k], tg[k], ou[k]) for k in range(sets)], reps=20)
    out["pd_1024x12"]["note"] = "launch floor: L2-resident, 12 CTAs"
    del st, tg, ou

    # P at the headline size WITHOUT the fused statistics epilogue (the headline loop accumulates them): what the
    # statistics cost, and how close the plain law runs to the measured copy bandwidth
    n, sets = 1_048_576, 4
    pi = syn.pd_inputs(n, NUM_DOFS, seed=1)
    c4 = PDController(NUM_DOFS, pi.kp, pi.kd, tau_max=pi.tau_max, device=device)
    st = [pi.dof_state.to(device).clone() for _ in range(sets)]
    tg = [pi.q_target.to(device).clone() for _ in range(sets)]
    ou = [torch.empty(n, NUM_DOFS, device=device) for _ in range(sets)]
    record("pd_1048576x12_nostats", n, PD_BYTES_PER_ENV, [c4.bind(st[k], tg[k], ou[k]) for k in range(sets)], reps=20)
    del st, tg, ou

    # P on the Franka's own DOF count (D = 9: a 128-bit vector straddles two envs, per-element DOF indices)
    n, sets, d9 = 1_048_576, 4, 9
    pi = syn.pd_inputs(n, d9, seed=1)
    c9 = PDController(d9, pi.kp, pi.kd, tau_max=pi.tau_max, device=device)
    st = [pi.dof_state.to(device).clone() for _ in range(sets)]
    tg = [pi.q_target.to(device).clone() for _ in range(sets)]
    ou = [torch.empty(n, d9, device=device) for _ in range(sets)]
    record("pd_1048576x9_nostats", n, 16 * d9, [c9.bind(st[k], tg[k], ou[k]) for k in range(sets)], reps=20)
    del st, tg, ou

    # S fused step at C2 (6.8 MB per set) and at 1M envs (109 MB per set), both precisions; "_stats" = with the
    # statistics vector the rollout harness passes (persistent CTAs, one commit per CTA)
    from test_isaacgym_b200 import _lib
    sbuf = _lib.stats_buffer(device)
    for n, sets, reps in ((65_536, 24, 10), (1_048_576, 3, 20)):
        base = syn.servo_root_state(n, seed=2).to(device)
        bufs = [base.clone() for _ in range(sets)]
        for tag, prec in (("ref", 0), ("fast", PRECISION_FAST)):
            step = ServoStep(1600, 900, precision=prec)
            record(f"servo_step_{tag}_{n}", n, SERVO_BYTES_PER_ENV, [step.bind(b) for b in bufs], reps=reps)
            record(f"servo_step_{tag}_{n}_stats", n, SERVO_BYTES_PER_ENV, [step.bind(b, stats=sbuf) for b in bufs], reps=reps)
        del bufs, base

    # O at C3 (16,384 envs: 47 MB of gym tensors per set) and at 262,144 envs (750 MB per set)
    for n, sets, reps in ((16_384, 4, 20), (262_144, 2, 20)):
        fi = syn.franka_inputs(n, seed=3)
        for prec, ptag in ((0, "fp64chain"), (1, "fp32")):
            osc_calls, ik_calls, keep = [], [], []
            for _ in range(sets):
                d = fi.__class__(**{k: (v.to(device).clone() if isinstance(v, torch.Tensor) else v) for k, v in fi.__dict__.items()})
                o = torch.zeros(n, 9, device=device)
                ctl.bind(j_eef=d.j_eef, mm=d.mm, dof_pos=d.dof_pos, dof_vel=d.dof_vel,
                         default_dof_pos_tensor=d.default_dof_pos, num_envs=n, precision=prec)
                ctl.bind_hand(d.rb_states, d.hand_idxs)
                osc_calls.append(ctl.bind_control_osc(d.dpose, o[:, :7]))
                ik_calls.append(ctl.bind_control_ik(d.dpose, o[:, :7], dof_pos=d.dof_pos))
                keep.append((d, o))
            record(f"osc_{ptag}_{n}", n, OSC_BYTES_PER_ENV, osc_calls, reps, flops=OSC_FLOPS)
            record(f"ik_{ptag}_{n}", n, IK_BYTES_PER_ENV, ik_calls, reps, flops=IK_FLOPS)
            del osc_calls, ik_calls, keep
        ctl.bind(precision=0)

    # the whole pick step of examples/franka_cube_ik_osc.py:348-410 at C3: goal logic + OSC as two kernels
    n, sets = 16_384, 4
    ti, fi = syn.franka_task_inputs(n, seed=4), syn.franka_inputs(n, seed=5)
    task_calls, step_calls, keep = [], [], []
    for _ in range(sets):
        t = ti.__class__(**{k: (v.to(device).clone() if isinstance(v, torch.Tensor) else v) for k, v in ti.__dict__.items()})
        d = fi.__class__(**{k: (v.to(device).clone() if isinstance(v, torch.Tensor) else v) for k, v in fi.__dict__.items()})
        dpose = torch.zeros(n, 6, 1, device=device)
        pos_action, effort = torch.zeros(n, 9, device=device), torch.zeros(n, 9, device=device)
        task = ctl.TaskStep(t.rb_states, t.box_idxs, t.hand_idxs, t.dof_pos, t.init_pos, t.init_rot, t.hand_restart, "osc")
        ctl.bind(j_eef=d.j_eef, mm=d.mm, dof_pos=t.dof_pos, dof_vel=t.dof_state[:, 1].view(n, 9, 1),
                 default_dof_pos_tensor=d.default_dof_pos, num_envs=n, precision=0)
        ctl.bind_hand(t.rb_states, t.hand_idxs)
        tc = task.bind(dpose, pos_action[:, 7:9])
        oc = ctl.bind_control_osc(dpose, effort[:, :7])
        task_calls.append(tc)
        step_calls += [tc, oc]
        keep.append((t, d, dpose, pos_action, effort, task))
    record(f"franka_task_{n}", n, 150, task_calls, 20)
    ms_pair = graph_time(step_calls, device, 20) * 2      # per (task + osc) pair
    out[f"franka_pick_step_{n}"] = {"envs": n, "us_per_step": round(ms_pair * 1e3, 3), "env_steps_per_s": n / (ms_pair * 1e-3),
                                    "kernels_per_step": 2, "note": "franka_task + osc (fp64 chain), CUDA-graph replay"}
    fused = []
    for (t, d, dpose, pos_action, effort, task) in keep:
        ctl.bind(j_eef=d.j_eef, mm=d.mm, dof_pos=t.dof_pos, dof_vel=t.dof_state[:, 1].view(n, 9, 1),
                 default_dof_pos_tensor=d.default_dof_pos, num_envs=n, precision=0)
        fused.append(ctl.bind_pick_osc(task, effort[:, :7], pos_action[:, 7:9]))
    ms_fused = graph_time(fused, device, 20)
    out[f"franka_pick_step_fused_{n}"] = {"envs": n, "us_per_step": round(ms_fused * 1e3, 3), "env_steps_per_s": n / (ms_fused * 1e-3),
                                          "kernels_per_step": 1, "note": "b200ctl_franka_pick_osc (fp64 chain), CUDA-graph replay"}
    fused_ik = []
    for (t, d, dpose, pos_action, effort, task) in keep:
        ctl.bind(j_eef=d.j_eef, num_envs=n, precision=0)
        fused_ik.append(ctl.bind_pick_ik(task, pos_action[:, :7], pos_action[:, 7:9]))
    ms_ik = graph_time(fused_ik, device, 20)
    out[f"franka_pick_ik_step_fused_{n}"] = {"envs": n, "us_per_step": round(ms_ik * 1e3, 3), "env_steps_per_s": n / (ms_ik * 1e-3),
                                             "kernels_per_step": 1, "note": "b200ctl_franka_pick_ik (default controller, fp64 chain), CUDA-graph replay"}
    return out


# ------------------------------------------------------------------------------------------ CPU baseline / reference arm
def cpu_pd_step_fn(num_envs, seed=0):
    """The oracle port: torch-CPU evaluation of the reference's joint-PD expression with all host threads."""
    from oracle import pd as opd
    from test_isaacgym_b200 import synthetic as syn
    pi = syn.pd_inputs(num_envs, NUM_DOFS, seed=seed, gain_set="B")
    torch.set_num_threads(os.cpu_count() or 1)
    return lambda: opd.pd_torque(pi.dof_state, pi.q_target, pi.kp, pi.kd, tau_max=pi.tau_max)


def cpu_baseline(budget_s=12.0):
    fn = cpu_pd_step_fn(ENVS_PER_GPU)
    fn()
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < 5 or (time.perf_counter() < t_end and len(times) < 40):
        t0 = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return {"value": ENVS_PER_GPU / med, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{len(times)} steps of the full {ENVS_PER_GPU} x {NUM_DOFS} workload, median {med * 1e3:.1f} ms/step, "
                      "oracle/pd.py torch-CPU fp32 (the reference has no function for this law; no _ref build)"}


def cpu_family_baselines(families):
    """CPU oracle (the reference's numpy / scipy / torch ops, all host threads) for the other families at their
    BASELINE config sizes, timed in the same run: adds `cpu_env_steps_per_s` / `speedup_vs_cpu` to the entries."""
    from oracle import servo as osv, franka as ofr
    from test_isaacgym_b200 import synthetic as syn
    import numpy as np
    torch.set_num_threads(os.cpu_count() or 1)

    def timed(fn, reps):
        fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return statistics.median(ts)

    n = 65_536
    state = syn.servo_root_state(n, seed=2)
    t = timed(lambda: osv.servo_step(state, 1600, 900), 3)
    for key in (f"servo_step_ref_{n}", f"servo_step_fast_{n}"):
        if key in families:
            families[key]["cpu_env_steps_per_s"] = n / t
            families[key]["speedup_vs_cpu"] = families[key]["env_steps_per_s"] / (n / t)
    n = 16_384
    fi = syn.franka_inputs(n, seed=3)
    kp, kpn = 150.0, 10.0
    t_osc = timed(lambda: ofr.control_osc(fi.dpose, fi.j_eef, fi.mm, fi.dof_pos, fi.dof_vel, fi.hand_vel, fi.default_dof_pos,
                                          kp, 2 * np.sqrt(kp), kpn, 2 * np.sqrt(kpn)), 10)
    t_ik = timed(lambda: ofr.control_ik(fi.dpose, fi.j_eef, 0.05), 10)
    for key, t in ((f"osc_fp64chain_{n}", t_osc), (f"osc_fp32_{n}", t_osc), (f"ik_fp64chain_{n}", t_ik), (f"ik_fp32_{n}", t_ik)):
        if key in families:
            families[key]["cpu_env_steps_per_s"] = n / t
            families[key]["speedup_vs_cpu"] = families[key]["env_steps_per_s"] / (n / t)
    families["_cpu_note"] = (f"cpu_env_steps_per_s: oracle (reference arithmetic: numpy/scipy fp64 for servo, torch fp32 for "
                             f"osc/ik) on {torch.get_num_threads()} host threads, median of 3-10 steps at the same size")


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    n = ENVS_PER_GPU
    fn = cpu_pd_step_fn(n)
    t0 = time.perf_counter()
    fn()
    t_one = time.perf_counter() - t0
    # bound the run: shrink the per-step sample if K full-size steps would take more than ~100 s
    total = (args.steps + args.warmup) * t_one
    sample = n
    if total > 100.0:
        sample = max(65_536, int(n * 100.0 / total))
        fn = cpu_pd_step_fn(sample)
    for _ in range(args.warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3 * (n / sample), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(), "sample_envs_per_step": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{args.steps} steps x {sample} envs x {NUM_DOFS} DOF (oracle/pd.py, torch-CPU fp32, "
                                       f"{torch.get_num_threads()} threads); the reference has no single function for this law"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------ our arm
def run_b200(args):
    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        emit({"error": "no CUDA device: b200ctl has no CPU path"})
        return 1
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    import torch.distributed as dist
    if world > 1:
        # NCCL on a HIGH-PRIORITY stream: the control kernels are persistent grids chained by programmatic dependent
        # launch, so a normal-priority collective kernel is starved of an SM slot until the control stream itself
        # blocks on the window's event -- and then waits for the peer to reach the same point (measured at N=2:
        # 47.8 us/step instead of 32.6)
        from test_isaacgym_b200.sharding import nccl_options
        dist.init_process_group("nccl", device_id=device, pg_options=nccl_options())
    from test_isaacgym_b200 import _lib
    from test_isaacgym_b200.sharding import StatsReducer, StatsWindow, env_slice
    from test_isaacgym_b200.pd_control import pd_torque

    peak, peak_src = hbm_peak()
    total_envs = ENVS_PER_GPU * world
    lo, hi = env_slice(total_envs, rank, world)
    wl = PdWorkload(device, hi - lo, seed=1000 + rank)
    reducer = StatsReducer("torch", device) if world > 1 else None
    stats_every = max(1, args.stats_every)
    wl.bind(StatsWindow(device, reducer, stats_every, overlap=args.stats_overlap))
    step = wl.step

    with ClockSampler(local_rank) as clocks:
        # sustained warm-up so the clock samples describe the loaded state of this very kernel
        # (every rank must run the SAME number of steps: the statistics all-reduce fires every k-th step, and ranks
        # whose clocks disagree on when the warm-up ends would enqueue different numbers of collectives and deadlock;
        # the continue / stop decision is therefore itself reduced over the ranks)
        t_end = time.perf_counter() + args.sustain_s
        it = 0
        while True:
            go = torch.tensor([1 if time.perf_counter() < t_end else 0], device=device, dtype=torch.int32)
            if world > 1:
                dist.all_reduce(go, op=dist.ReduceOp.MIN)
            if int(go.item()) == 0:
                break
            for _ in range(200):
                wl.step(it)
                it += 1
            torch.cuda.synchronize(device)
        for i in range(args.warmup):
            step(i)
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        launches0 = _lib.launch_count()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(device)
        start.record()
        for i in range(args.steps):
            step(i)
        end.record()
        torch.cuda.synchronize(device)
        launches = _lib.launch_count() - launches0
        wl.window.finish()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([start.elapsed_time(end)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms_total = ms.item()

        # ---- e2e: public API on pinned HOST buffers, H2D + kernel + D2H inside the timed region
        h = wl.host
        hs, ht = h.dof_state.pin_memory(), h.q_target.pin_memory()
        hout = torch.empty(hi - lo, NUM_DOFS, dtype=torch.float32, pin_memory=True)
        e2e_steps = max(3, min(args.steps, 20))
        for _ in range(3):
            pd_torque(hs, ht, h.kp, h.kd, tau_max=h.tau_max, out=hout)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            pd_torque(hs, ht, h.kp, h.kd, tau_max=h.tau_max, out=hout)
        e2e_s = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e_checksum = float(hout[::4097].double().abs().sum())

    ms_per_step = ms_total / args.steps
    value = total_envs * args.steps / (ms_total * 1e-3)
    per_gpu_rate = (hi - lo) / (ms_per_step * 1e-3)
    achieved = per_gpu_rate * PD_BYTES_PER_ENV / 1e9
    e2e_value = total_envs * e2e_steps / e2e_s.item()

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(), "envs_per_gpu": hi - lo, "num_dofs": NUM_DOFS, "global_envs": total_envs,
                   "parallelism": f"env-slices x{world}", "stats_allreduce_every": stats_every if world > 1 else None,
                   "stats_allreduce": ("side stream" if args.stats_overlap else "in order on the control stream") if world > 1 else None,
                   "l2_policy": f"inputs > L2: {wl.sets} rotating buffer sets of 201 MB (151 MB in + 50 MB out each)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "pd_torque_vec4_kernel", "bytes_per_launch": (hi - lo) * PD_BYTES_PER_ENV,
                     "peak_source": peak_src, "of": "measured" if peak_src.startswith("measured") else "fallback"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": (hs.numel() + ht.numel()) * 4 * world,
                "d2h_bytes_per_step": hout.numel() * 4 * world, "steps": e2e_steps,
                "api": "test_isaacgym_b200.pd_control.pd_torque(host tensors) -> b200ctl_pd_torque_host",
                "checksum": e2e_checksum},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
    }
    traffic_file = os.path.join(ROOT, "profiles", "pd_traffic.json")
    if os.path.isfile(traffic_file):
        try:
            line["roofline"]["traffic"] = json.load(open(traffic_file))["dram_bytes_per_launch"]
        except Exception:
            pass
    if rank == 0 and world == 1 and not args.no_families:
        # give the headline's 1.6 GB of buffers back first: the family entries are then laid out in device memory as
        # in a process of their own (with the buffers alive the 262,144-env OSC entry read 53.9 us instead of 50.9)
        del step, wl, hs, ht, hout, h
        torch.cuda.empty_cache()
        line["families"] = family_numbers(device, peak)
    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline()
        if "families" in line:
            cpu_family_baselines(line["families"])
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_JSON_FD = None


def claim_stdout():
    """Keep file descriptor 1 for the JSON line alone: native libraries write there too (NCCL prints its version banner
    to stdout under NCCL_DEBUG=VERSION, ahead of the line), so fd 1 is pointed at stderr for the rest of the run and the
    line goes to a duplicate of the original."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
        return
    while data:
        data = data[os.write(_JSON_FD, data):]


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--stats-every", type=int, default=16, help="all-reduce the statistics vector every k steps (N > 1)")
    ap.add_argument("--stats-overlap", action="store_true", help="all-reduce on a side stream instead of in order (N > 1)")
    ap.add_argument("--sustain-s", type=float, default=1.0, help="seconds of pre-load before the timed region (clock sampling)")
    ap.add_argument("--no-families", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    return run_reference(args) if args.impl == "reference" else run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
