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
vector is exchanged EVERY step inside the PD kernel itself (--stats-collective fused, the default:
one publisher CTA per launch stores the previous step's vector into every rank's mailbox over
NVLink peer memory and sums the rows of the step before).  Other forms: the library's stand-alone
all-reduce kernel (peer / peer-lagged, every --stats-every steps) or NCCL (nccl).  N > 1 lines
also carry `strong` (1,048,576 envs IN TOTAL split over the N GPUs) and `per_step_stats` (the same
loop under the other exchange forms, measured side by side in this process).  The loop is issued the
way the rollout harness issues control steps: consecutive steps over the rotating buffer sets captured
once as CUDA graphs (20 steps, and 4 for what does not fill 20) and replayed (`step_issue`; --no-graph
issues single bound calls).  The timed region is exactly K steps between two CUDA events on the launching
stream, bracketed by barrier + synchronize; one untimed buffer rotation is enqueued between the
synchronize and the start event so that the region opens on a busy stream (`timed_region`, which also
reports the same K steps opened on an idle device; --no-lead-in makes that the headline).
If the peer mailboxes cannot be mapped (no peer access) every rank falls back to NCCL in order and
the line says so (`stats_collective_note`).

Printed keys beyond the base contract:
  roofline      dominant kernel vs measured HBM peak (MEASURED_PEAKS.json), algorithmic bytes; `copy_here` = the peak's
                own measurement (torch copy, burst / sustained) repeated in this process over zeros and over random bits
                with SM clock / board power: the sustained loop runs at the board's power cap, and power follows the data
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
IK_BYTES_PER_ENV, IK_FLOPS, IK_FLOPS_LITERAL = 248, 480, 1530        # canonical (Cholesky route) / literal
OSC_BYTES_PER_ENV, OSC_FLOPS, OSC_FLOPS_LITERAL = 496, 1850, 5270     # (the reference's LU inversions + dense bmm)
L2_BYTES = 126 << 20
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


def copy_bandwidth_here(device, index, seconds=1.0):
    """The measurement behind MEASURED_PEAKS.json's hbm_gbs, repeated on THIS box in THIS process: `b.copy_(a)` over 1 Gi
    bf16 elements (read + write bytes), best of 10 single launches (burst) and back to back for `seconds` (sustained: the
    state the headline loop runs in), once over ZEROS and once over RANDOM BITS, with SM clock / board power / throttle
    reasons sampled during each sustained run.  The board's power depends on the data: the same copy draws ~660-790 W
    over zeros and sits at the 1,000 W cap over random bits (profiles/r02_pd_power.txt) -- the PD loop streams real floats.
    Informational: `roofline.frac` stays against the driver-written peak."""
    n = 1 << 30
    b = torch.empty(n, dtype=torch.bfloat16, device=device)
    nbytes = 2.0 * n * 2
    ev = lambda: torch.cuda.Event(enable_timing=True)
    out = {"how": "torch b.copy_(a) over 1 Gi bf16 elements, read + write bytes: best of 10 launches (burst) / back to back for "
                  f"{seconds} s (sustained), this box, this process, right after the headline loop; source = zeros / random bits"}
    for tag in ("zeros", "random"):
        a = (torch.zeros(n, dtype=torch.bfloat16, device=device) if tag == "zeros" else
             torch.randint(-32768, 32767, (n,), dtype=torch.int16, device=device).view(torch.bfloat16))
        for _ in range(3):
            b.copy_(a)
        torch.cuda.synchronize(device)
        best = 1e9
        for _ in range(10):
            s, e = ev(), ev()
            s.record(); b.copy_(a); e.record()
            torch.cuda.synchronize(device)
            best = min(best, s.elapsed_time(e))
        with ClockSampler(index) as smp:
            reps, t0 = 0, time.perf_counter()
            s, e = ev(), ev()
            s.record()
            while time.perf_counter() - t0 < seconds:
                for _ in range(32):
                    b.copy_(a)
                reps += 32
                if reps % 256 == 0:
                    torch.cuda.synchronize(device)      # bound the launch queue
            e.record()
            torch.cuda.synchronize(device)
        c = smp.summary()
        out[tag] = {"burst_gbs": nbytes / best / 1e6, "sustained_gbs": nbytes / (s.elapsed_time(e) / reps) / 1e6,
                    "sm_mhz": c.get("sm_mhz"), "board_w": c.get("board_w"), "reasons": c.get("reasons")}
        del a
    del b
    torch.cuda.empty_cache()
    return out


def workload_name():
    return f"pd_torque: {ENVS_PER_GPU} envs x {NUM_DOFS} DOF per GPU (BASELINE configs[3], one env slice per GPU)"


def make_config(world, envs_per_gpu, stats_every=None, stats_mode=None, sets=4):
    """`config` of the JSON line -- the SAME keys and values for the repo arm and the reference arm, so the two lines
    describe one workload (the driver compares them)."""
    return {"workload": workload_name(), "envs_per_gpu": envs_per_gpu, "num_dofs": NUM_DOFS, "global_envs": envs_per_gpu * world,
            "parallelism": f"env-slices x{world}", "stats_allreduce_every": stats_every if world > 1 else None,
            "stats_allreduce": stats_mode if world > 1 else None,
            "l2_policy": f"inputs > L2: {sets} rotating buffer sets of 201 MB (151 MB in + 50 MB out each)"}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the benchmark runs."""

    def __init__(self, index: int, period_s: float = 0.02):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.watts = [], set(), []
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
                try:
                    self.watts.append(nv.nvmlDeviceGetPowerUsage(self._h) / 1e3)
                except Exception:
                    pass
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
        out = {"sm_mhz": statistics.median(busy), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
               "samples": len(self.samples)}
        if self.watts:
            out["board_w"] = statistics.median(self.watts[len(self.watts) // 2:])      # settled half of the run
        return out


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


class StepLoop:
    """The headline loop.  `sets` consecutive PD steps (one per rotating buffer set) are captured ONCE as a CUDA graph
    (`StepGraph`, the way the rollout harness issues its control steps) and replayed; steps that do not fill a whole
    block are issued as single bound calls.  The statistics window logic (buffer swap, all-reduce every k steps) stays
    outside the graph on the same stream.  Measured on one B200 (profiles/r02_k20_probe.txt): a 20-step timed region
    reads 33.9-34.1 us/step issued call by call (the first launch after the synchronize arrives ~10 us late), 33.4 at
    400 steps, and 33.0-33.3 as graph replays (kernel-to-kernel edges inside a graph are tighter than stream launches)."""

    LONG = 20       # steps per long graph: the timed region of a default driver run (--steps 20) is one replay

    def __init__(self, device, wl, fused=None, use_graph=True):
        from test_isaacgym_b200.graph import StepGraph
        self.wl, self.fused, self.sets = wl, fused, wl.sets
        self.pos = 0            # steps issued so far; pos % sets is the buffer set of the next step
        self.replays = 0        # graph replays issued
        self.graph_steps = 0    # steps issued through graph replays
        for _ in range(self.sets):      # every entry point has run once before capture (module load)
            self._single()
        # One graph per rotation offset r (steps over sets r, r+1, ... mod sets), so that a timed region can start anywhere,
        # in two lengths: `sets` steps, and the multiple of `sets` nearest below LONG steps (graph-to-graph boundaries cost
        # ~1.2 us each, profiles/r02_k20_probe.txt: 33.4 us/step in 4-step graphs, 33.1 in one 20-step graph).  A block
        # never straddles a statistics exchange that is issued between steps (window forms), so those cap at `every`.
        self.graphs = self.long_graphs = None
        self.block = self.sets
        window_form = fused is None and wl.window.reducer is not None
        cap = min(self.LONG, wl.window.every) if window_form else self.LONG
        if not use_graph or cap < self.sets:
            return
        self.block = self.sets * (cap // self.sets)
        calls_of = (lambda k, b: fused[k]) if fused is not None else (lambda k, b: wl.calls[k][b])
        banks = (0,) if fused is not None else (0, 1)    # fused: accumulator parity = set parity (sets is even)

        def capture(length):
            return [[StepGraph([calls_of((r + j) % self.sets, b) for j in range(length)], device, warmup=0) for b in banks]
                    for r in range(self.sets)]
        self.graphs = capture(self.sets)
        if self.block > self.sets:
            self.long_graphs = capture(self.block)

    def _single(self):
        if self.fused is not None:
            self.fused[self.pos % self.sets]()
        else:
            self.wl.step(self.pos)
        self.pos += 1

    def _fits(self, length):
        """A replayed block must not straddle a statistics exchange (the all-reduce is issued between two steps).  With no
        reducer (N = 1) nothing is exchanged: the local accumulator swap simply happens after the block."""
        if self.fused is not None or self.wl.window.reducer is None:
            return True
        w = self.wl.window
        return w.every - (w._steps % w.every) >= length

    def _replay(self, graphs, length):
        r = self.pos % self.sets
        if self.fused is not None:
            graphs[r][0]()
        else:
            graphs[r][self.wl.window.cur]()
            for _ in range(length):
                self.wl.window.step_done()
        self.replays += 1
        self.graph_steps += length
        self.pos += length

    def run(self, k):
        done = 0
        while done < k:
            if self.long_graphs is not None and k - done >= self.block and self._fits(self.block):
                self._replay(self.long_graphs, self.block)
                done += self.block
            elif self.graphs is not None and k - done >= self.sets and self._fits(self.sets):
                self._replay(self.graphs, self.sets)
                done += self.sets
            else:
                self._single()
                done += 1


# ------------------------------------------------------------------------------------------ families (rank 0, N=1)
def graph_time(calls, device, reps, warm=3, runs=5, warm_ms=30.0, stat="min"):
    """Capture `calls` (bound C-ABI calls, one kernel each) into one CUDA graph (`StepGraph`) and time `reps`
    replays with CUDA events on the replaying stream, `runs` times.  Returns ms per kernel launch: the best run
    (`stat="min"`), the median (`"median"`) or both (`"both"` -> (best, median)).  The graph removes the host launch
    cost, which at 64K envs is larger than the kernels themselves.  Replays run for `warm_ms` before the first timed
    run: these entries follow seconds of host-side input generation, and an idle GPU needs tens of milliseconds to come
    back to its boost clock (a single timed run after three warm replays read 5-8 % high on the latency-bound kernels)."""
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
    best, med = times[0], times[len(times) // 2]
    return (best, med) if stat == "both" else (best if stat == "min" else med)


def sets_for(touched_bytes_per_set, lo=3, hi=64):
    """Rotating buffer sets needed for the bytes a launch actually TOUCHES (not the bytes allocated: the O kernels read
    15 MB of a 47 MB set at 16,384 envs) to exceed twice the 126 MB L2 between two uses of the same set."""
    return int(max(lo, min(hi, -(-2 * L2_BYTES // max(1, touched_bytes_per_set)))))


def _lib_stats(device):
    from test_isaacgym_b200 import _lib
    return _lib.stats_buffer(device)


def measure_fma_peaks(device):
    """FMA-pipe peaks measured by the library's own micro-benchmark in this process (MEASURED_PEAKS.json has none)."""
    from test_isaacgym_b200 import _lib
    f32b, f32m = _lib.measure_fma_peak("f32", device, 7)
    f64b, f64m = _lib.measure_fma_peak("f64", device, 7)
    return {"fp32_tflops": f32b, "fp32_tflops_median": f32m, "fp64_tflops": f64b, "fp64_tflops_median": f64m,
            "how": "b200ctl_measure_fma_peak: 16 dependent-FMA chains per thread, full occupancy, 7 launches of ~5 ms, CUDA events; "
                   "nominal 148 SM x 128 (64) lanes x 2 x 1.965 GHz = 74.4 (37.2) TFLOP/s"}


def family_numbers(device, peak_gbs):
    """Per-GPU throughput of the other laws / sizes named in BASELINE.json configs (not bench lines of their own).
    Every entry rotates enough buffer sets to exceed the 126 MB L2 and replays a CUDA graph of bound calls."""
    from test_isaacgym_b200 import synthetic as syn
    from test_isaacgym_b200.pd_control import PDController
    from test_isaacgym_b200.servo_step import ServoStep, PRECISION_FAST
    import test_isaacgym_b200.franka_cube_ik_osc as ctl
    out = {"_timing_note": "each entry: CUDA-graph replays of bound calls over rotating buffer sets whose TOUCHED bytes exceed "
                           "twice the 126 MB L2, 30 ms of warm replays, 5 timed runs of `reps` replays: us_per_step = best, "
                           "us_per_step_median = median"}

    fma = out["_fma_peak"] = measure_fma_peaks(device)

    def record(name, n, bytes_per_env, calls, reps, flops=None, flops_literal=None, fp64_chain=False):
        ms, ms_med = graph_time(calls, device, reps, stat="both")
        rate = n / (ms * 1e-3)
        e = {"envs": n, "us_per_step": round(ms * 1e3, 3), "us_per_step_median": round(ms_med * 1e3, 3), "env_steps_per_s": rate,
             "hbm_frac": rate * bytes_per_env / (peak_gbs * 1e9), "buffer_sets": len(calls)}
        if flops:
            # FP32-pipe utilisation the metric asks for, against the MEASURED FMA peak: canonical = the flops of the
            # factorisation route this kernel takes, literal = the reference's own operation count (LU inversions, bmm)
            e["tflops_canonical"] = rate * flops / 1e12
            e["fp32_frac_canonical"] = rate * flops / 1e12 / fma["fp32_tflops"]
            e["fp32_frac_literal"] = rate * flops_literal / 1e12 / fma["fp32_tflops"]
            if fp64_chain:      # the default precision runs the chain on the fp64 pipe: that is the pipe it occupies
                e["fp64_pipe_frac"] = rate * flops / 1e12 / fma["fp64_tflops"]
        # DRAM bytes per env actually moved (ncu dram__bytes_read + write at the throughput size, profiles/): the gym
        # layouts (13-float rows, a 6x7 slot of a 10x6x9 jacobian, a 7x7 corner of a 9x9 matrix) are fetched at
        # 64-byte granularity, so this exceeds the algorithmic bytes without any re-read
        for key, (b, src) in DRAM_BYTES_PER_ENV.items():
            if name.startswith(key):
                e["dram_bytes_per_env_ncu"] = b
                e["dram_frac"] = rate * b / (peak_gbs * 1e9)
                e["dram_bytes_source"] = src
                e["touched_mb_per_set"] = round(n * b / 1e6, 1)
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
    # the fully featured instantiation (north_star: "fused clamp/limit/gain logic in one pass"): floor-mod angle wrap +
    # target clamp into the joint limits + a velocity-target tensor + torque saturation, 240 B/env; with and without
    # the statistics epilogue
    cf = PDController(NUM_DOFS, pi.kp, pi.kd, tau_max=pi.tau_max, q_lo=pi.q_lo, q_hi=pi.q_hi, wrap_angle=True,
                      clamp_target=True, device=device)
    qd = [syn.pd_inputs(n, NUM_DOFS, seed=2, qd_target_std=1.0).qd_target.to(device) for _ in range(sets)]
    full_bytes = PD_BYTES_PER_ENV + 4 * NUM_DOFS
    record("pd_1048576x12_fullflags", n, full_bytes, [cf.bind(st[k], tg[k], ou[k], qd_target=qd[k]) for k in range(sets)], reps=20)
    fs = _lib_stats(device)
    record("pd_1048576x12_fullflags_stats", n, full_bytes,
           [cf.bind(st[k], tg[k], ou[k], qd_target=qd[k], stats=fs) for k in range(sets)], reps=20)
    for k_ in ("pd_1048576x12_fullflags", "pd_1048576x12_fullflags_stats"):
        out[k_]["kernel"] = "pd_torque_vec4_kernel<WRAP=1,CLAMP_TGT=1,HAS_QD=1,HAS_TMAX=1>"
    del st, tg, ou, qd

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

    # O at C3 (16,384 envs: 47 MB of gym tensors per set, of which a launch touches 15.7 MB (OSC) / 7.5 MB (IK)) and at
    # 262,144 envs (750 MB per set).  The number of rotating sets follows the TOUCHED bytes (IK's, the smaller): round 1
    # rotated 4 sets at C3 -- 188 MB allocated but only 61 MB touched per cycle, i.e. L2-resident.
    def to_dev(obj):
        return obj.__class__(**{k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in obj.__dict__.items()})

    def dev_clone(obj):
        return obj.__class__(**{k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in obj.__dict__.items()})

    for n, reps in ((16_384, 10), (262_144, 20)):
        sets = sets_for(n * DRAM_BYTES_PER_ENV["ik_"][0], lo=3, hi=40)
        fi = syn.franka_inputs(n, seed=3)
        base = to_dev(fi)
        keep = [base] + [dev_clone(base) for _ in range(sets - 1)]
        outs = [torch.zeros(n, 9, device=device) for _ in range(sets)]
        for prec, ptag in ((0, "fp64chain"), (1, "fp32")):
            osc_calls, ik_calls = [], []
            for d, o in zip(keep, outs):
                ctl.bind(j_eef=d.j_eef, mm=d.mm, dof_pos=d.dof_pos, dof_vel=d.dof_vel,
                         default_dof_pos_tensor=d.default_dof_pos, num_envs=n, precision=prec)
                ctl.bind_hand(d.rb_states, d.hand_idxs)
                osc_calls.append(ctl.bind_control_osc(d.dpose, o[:, :7]))
                ik_calls.append(ctl.bind_control_ik(d.dpose, o[:, :7], dof_pos=d.dof_pos))
            record(f"osc_{ptag}_{n}", n, OSC_BYTES_PER_ENV, osc_calls, reps, flops=OSC_FLOPS, flops_literal=OSC_FLOPS_LITERAL,
                   fp64_chain=prec == 0)
            record(f"ik_{ptag}_{n}", n, IK_BYTES_PER_ENV, ik_calls, reps, flops=IK_FLOPS, flops_literal=IK_FLOPS_LITERAL,
                   fp64_chain=prec == 0)
            if prec == 1:
                for k_ in (f"osc_{ptag}_{n}", f"ik_{ptag}_{n}"):
                    out[k_]["tolerance_note"] = ("all-fp32 chain: meets the 1e-4 bar only on well-conditioned envs (tests/test_gpu_franka.py: "
                                                 "bound stated and asserted there); the fp64-chain entries are the parity-bar numbers")
            del osc_calls, ik_calls
        del keep, outs, base
        ctl.bind(precision=0)
        torch.cuda.empty_cache()

    # the whole pick step of examples/franka_cube_ik_osc.py:348-410 at C3: goal logic + OSC as two kernels
    n = 16_384
    sets = sets_for(n * DRAM_BYTES_PER_ENV["franka_task"][0], lo=4, hi=40)
    ti, fi = syn.franka_task_inputs(n, seed=4), syn.franka_inputs(n, seed=5)
    t0_, d0_ = to_dev(ti), to_dev(fi)
    task_calls, step_calls, keep = [], [], []
    for k_ in range(sets):
        t = t0_ if k_ == 0 else dev_clone(t0_)
        d = d0_ if k_ == 0 else dev_clone(d0_)
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
    record(f"franka_task_{n}", n, 150, task_calls, 10)
    ms_pair = graph_time(step_calls, device, 10) * 2      # per (task + osc) pair
    out[f"franka_pick_step_{n}"] = {"envs": n, "us_per_step": round(ms_pair * 1e3, 3), "env_steps_per_s": n / (ms_pair * 1e-3),
                                    "kernels_per_step": 2, "note": "franka_task + osc (fp64 chain), CUDA-graph replay"}
    fused = []
    for (t, d, dpose, pos_action, effort, task) in keep:
        ctl.bind(j_eef=d.j_eef, mm=d.mm, dof_pos=t.dof_pos, dof_vel=t.dof_state[:, 1].view(n, 9, 1),
                 default_dof_pos_tensor=d.default_dof_pos, num_envs=n, precision=0)
        fused.append(ctl.bind_pick_osc(task, effort[:, :7], pos_action[:, 7:9]))
    ms_fused = graph_time(fused, device, 10)
    out[f"franka_pick_step_fused_{n}"] = {"envs": n, "us_per_step": round(ms_fused * 1e3, 3), "env_steps_per_s": n / (ms_fused * 1e-3),
                                          "kernels_per_step": 1, "note": "b200ctl_franka_pick_osc (fp64 chain), CUDA-graph replay"}
    fused_ik = []
    for (t, d, dpose, pos_action, effort, task) in keep:
        ctl.bind(j_eef=d.j_eef, num_envs=n, precision=0)
        fused_ik.append(ctl.bind_pick_ik(task, pos_action[:, :7], pos_action[:, 7:9]))
    ms_ik = graph_time(fused_ik, device, 10)
    out[f"franka_pick_ik_step_fused_{n}"] = {"envs": n, "us_per_step": round(ms_ik * 1e3, 3), "env_steps_per_s": n / (ms_ik * 1e-3),
                                             "kernels_per_step": 1, "note": "b200ctl_franka_pick_ik (default controller, fp64 chain), CUDA-graph replay"}
    # the loop law of examples/franka_osc.py:221-241 (all 9 DOFs) as one launch: hand-pose gather, quaternion
    # renormalisation, orientation error, dpose, OSC solve.  Algorithmic bytes: J 216 + M 324 + qd 36 + hand pose 28 +
    # pos_des 12 + orn_des 16 + u 36 = 668 B/env
    import test_isaacgym_b200.franka_osc as fosc
    osc9 = []
    for (t, d, dpose, pos_action, effort, task) in keep:
        pos_des = torch.zeros(n, 3, device=device)
        orn_des = torch.zeros(n, 4, device=device)
        orn_des[:, 3] = 1.0
        u9 = torch.zeros(n, 9, 1, device=device)
        osc9.append(fosc.bind_osc_step(t.rb_states, t.hand_idxs, pos_des, orn_des, d.jacobian[:, syn.FRANKA_JACOBIAN_SLOT],
                                       d.mass_matrix, t.dof_state[:, 1].view(n, 9, 1), u9))
        osc9[-1].keep = (osc9[-1].keep, pos_des, orn_des, u9)
    record(f"franka_osc_step_{n}", n, 668, osc9, 10)
    out[f"franka_osc_step_{n}"]["note"] = "b200ctl_franka_osc_step, 9 DOF, fp64 chain (examples/franka_osc.py:221-241 in one launch)"
    del keep, task_calls, step_calls, fused, fused_ik, osc9
    torch.cuda.empty_cache()

    # Small launches -- the reference scripts' own default is 256 envs.  north_star's "one warp (or a warp group) per env" form:
    # launches of at most 32 envs per SM run with EIGHT LANES per env (osc_lanes_kernel and its ik / pick twins: direct coalesced
    # global loads, three shared-memory meetings, redundant factorisations; bit-identical to the tile kernels), timed here next
    # to the one-thread-per-env tile kernel forced with b200ctl_osc_set_lanes(1)
    small = {"_note": "us per launch, fp64 chain, CUDA-graph replays over 40 rotating sets (L2-resident at these sizes whatever "
                      "the rotation); lanes8 = the form b200ctl picks by itself at this size, tile = one thread per env forced"}
    for n in (256, 4096):
        sets = 40
        t0_, d0_ = to_dev(syn.franka_task_inputs(n, seed=4)), to_dev(syn.franka_inputs(n, seed=5))
        keep = []
        for k_ in range(sets):
            t = t0_ if k_ == 0 else dev_clone(t0_)
            d = d0_ if k_ == 0 else dev_clone(d0_)
            keep.append((t, d, torch.zeros(n, 9, device=device), torch.zeros(n, 9, device=device),
                         ctl.TaskStep(t.rb_states, t.box_idxs, t.hand_idxs, t.dof_pos, t.init_pos, t.init_rot, t.hand_restart, "osc")))
        for form, mode in (("lanes8", -1), ("tile", 1)):
            _lib.osc_set_lanes(mode)
            calls = {"osc": [], "ik": [], "pick_osc": [], "pick_ik": []}
            for (t, d, pos_action, effort, task) in keep:
                ctl.bind(j_eef=d.j_eef, mm=d.mm, dof_pos=t.dof_pos, dof_vel=t.dof_state[:, 1].view(n, 9, 1),
                         default_dof_pos_tensor=d.default_dof_pos, num_envs=n, precision=0)
                ctl.bind_hand(t.rb_states, t.hand_idxs)
                calls["osc"].append(ctl.bind_control_osc(d.dpose, effort[:, :7]))
                calls["ik"].append(ctl.bind_control_ik(d.dpose, pos_action[:, :7], dof_pos=t.dof_pos))
                calls["pick_osc"].append(ctl.bind_pick_osc(task, effort[:, :7], pos_action[:, 7:9]))
                calls["pick_ik"].append(ctl.bind_pick_ik(task, pos_action[:, :7], pos_action[:, 7:9]))
            for k_, cs in calls.items():
                small.setdefault(f"{k_}_{n}", {})[f"{form}_us"] = round(graph_time(cs, device, 10) * 1e3, 3)
        _lib.osc_set_lanes(-1)
        del keep
    out["small_launch_forms"] = small
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


def stats_mode_name(args):
    if args.stats_collective == "fused":
        return ("EVERY step, inside the PD kernel (b200ctl_pd_torque_published): a publisher CTA stores the previous step's vector "
                "into every rank's mailbox over NVLink and sums the rows of the step before")
    where = "on a side stream next to the following step" if args.stats_overlap else "in order on the control stream"
    if args.stats_collective == "peer":
        return "b200ctl peer-memory all-reduce kernel (NVLink), " + where
    if args.stats_collective == "peer-lagged":
        return "b200ctl peer-memory all-reduce kernel (NVLink), one window lagged (never waits for a peer), " + where
    return "NCCL, " + ("side stream, one CTA slot reserved" if args.stats_overlap else "in order on the control stream")


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
            "config": make_config(world, ENVS_PER_GPU, 1 if args.stats_collective == "fused" else max(1, args.stats_every), stats_mode_name(args)),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample_envs_per_step": sample,
                             "sample": f"{args.steps} steps x {sample} envs x {NUM_DOFS} DOF (oracle/pd.py, torch-CPU fp32, "
                                       f"{torch.get_num_threads()} threads); the reference has no single function for this law"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------ our arm
def strong_scaling(device, rank, world, reducer, stats_every, overlap, steps, weak_ms_per_step, fused_pub=None):
    """BASELINE configs[3] as written: 1,048,576 envs IN TOTAL split as contiguous env slices over the N GPUs (strong
    scaling).  One CUDA graph holds one statistics window (`stats_every` PD steps over rotating buffer sets), the window's
    all-reduce follows it (in order, or on the side stream); device time by CUDA events, max over ranks.  Efficiency is
    against the single-GPU time of the same 1,048,576 envs measured in this run (the weak-scaling step of this rank)."""
    import torch.distributed as dist
    from test_isaacgym_b200 import _lib
    from test_isaacgym_b200.graph import StepGraph
    from test_isaacgym_b200.sharding import env_slice
    lo, hi = env_slice(ENVS_PER_GPU, rank, world)
    n = hi - lo
    sets = sets_for(n * PD_BYTES_PER_ENV, lo=4, hi=48)
    wl = PdWorkload(device, n, seed=2000 + rank, sets=sets)
    bufs = [_lib.stats_buffer(device), _lib.stats_buffer(device)]
    graphs = []
    from test_isaacgym_b200.sharding import PeerStatsReducer
    in_kernel = fused_pub is not None
    if in_kernel:
        # statistics exchanged every step inside the PD kernel (publisher CTA): one graph of an even number of steps
        stats_every = stats_every + (stats_every & 1)
        red = _lib.stats_buffer(device)
        calls = [wl.ctl.bind(wl.state[i % sets], wl.tgt[i % sets], wl.out[i % sets], stats=bufs[i & 1], stats_prev=bufs[(i & 1) ^ 1],
                             publish=fused_pub, reduced=red) for i in range(stats_every)]
        graphs = [StepGraph(calls, device, warmup=0)] * 2
    else:
        for b in bufs:
            calls = [wl.ctl.bind(wl.state[i % sets], wl.tgt[i % sets], wl.out[i % sets], stats=b) for i in range(stats_every)]
            graphs.append(StepGraph(calls, device, restore=[b]))
    events = [None, None]
    fused_zero = isinstance(reducer, PeerStatsReducer)

    def window(w):
        if in_kernel:
            graphs[0]()
            return
        k = w & 1
        if events[k] is not None:
            torch.cuda.current_stream(device).wait_event(events[k])
            events[k] = None
        if not fused_zero:
            bufs[k].zero_()
        graphs[k]()
        if fused_zero:
            reducer.all_reduce(bufs[k], zero_after=bufs[k ^ 1])
        elif reducer is not None:
            events[k] = reducer.all_reduce(bufs[k], overlap=overlap)

    windows = max(4, -(-steps // stats_every))
    for w in range(4):
        window(w)
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for w in range(windows):
        window(w)
    end.record()
    torch.cuda.synchronize(device)
    if reducer is not None:
        reducer.wait()
    ms = torch.tensor([start.elapsed_time(end)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    us = ms.item() * 1e3 / (windows * stats_every)
    ideal_us = weak_ms_per_step * 1e3 / world
    floor_us = 1.55 + n * PD_BYTES_PER_ENV / 6539.2e3        # one launch + the slice's bytes at the measured copy bandwidth
    return {"global_envs": ENVS_PER_GPU, "envs_per_gpu": n, "us_per_step": us, "env_steps_per_s": ENVS_PER_GPU / (us * 1e-6),
            "n1_us_per_step": weak_ms_per_step * 1e3, "efficiency_vs_n1": ideal_us / us,
            "stats_allreduce_every": 1 if in_kernel else stats_every,
            "buffer_sets": sets, "steps": windows * stats_every,
            "limiter": (f"{n} envs per GPU are {n * PD_BYTES_PER_ENV / 1e6:.1f} MB per step: {floor_us:.1f} us = one launch (1.55 us) + the "
                        f"bytes at HBM speed; the step no longer hides the launch floor (which is per launch, not per byte)")}


def host_link_ceiling(device, hs, ht, hout, world):
    """What the host link gives every rank when ALL ranks copy at once: the e2e step's 151 MB up and 50 MB down as plain
    async copies on two streams (no kernel), barrier-aligned, CUDA events.  The e2e figure cannot exceed this."""
    import torch.distributed as dist
    up, down = torch.cuda.Stream(device), torch.cuda.Stream(device)
    ds, dt_ = torch.empty_like(hs, device=device), torch.empty_like(ht, device=device)
    do = torch.empty(hout.shape, dtype=hout.dtype, device=device)
    best = 1e9
    for it in range(4):
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(up):
            ds.copy_(hs, non_blocking=True)
            dt_.copy_(ht, non_blocking=True)
        with torch.cuda.stream(down):
            hout.copy_(do, non_blocking=True)
        torch.cuda.synchronize(device)
        if it:
            best = min(best, time.perf_counter() - t0)
    t = torch.tensor([best], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def run_b200(args):
    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        emit({"error": "no CUDA device: b200ctl has no CPU path"})
        return 1
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    import torch.distributed as dist
    from test_isaacgym_b200 import _lib
    from test_isaacgym_b200.sharding import StatsReducer, StatsWindow, env_slice, nccl_options, configure_nccl_for_control_loops
    if world > 1:
        # NCCL sized to sit NEXT to the persistent control grids: one channel, 128-thread CTAs (= one CTA slot of the PD
        # kernel), on a high-priority stream; with --stats-overlap every persistent grid leaves that slot free
        configure_nccl_for_control_loops()
        dist.init_process_group("nccl", device_id=device, pg_options=nccl_options())
    from test_isaacgym_b200.pd_control import pd_torque

    peak, peak_src = hbm_peak()
    total_envs = ENVS_PER_GPU * world
    lo, hi = env_slice(total_envs, rank, world)
    wl = PdWorkload(device, hi - lo, seed=1000 + rank)
    from test_isaacgym_b200.sharding import PeerStatsReducer
    collective_note = None
    reducer = None
    if world > 1 and args.stats_collective != "nccl":
        try:
            reducer = PeerStatsReducer(device, lagged=args.stats_collective == "peer-lagged")
        except _lib.B200CtlError as e:      # raised on every rank together: no peer access on this box -> NCCL, in order
            collective_note = f"--stats-collective {args.stats_collective} unavailable ({e}); fell back to NCCL in order"
            args.stats_collective, args.stats_overlap = "nccl", False
    if world > 1 and reducer is None:
        reducer = StatsReducer("torch", device)
    in_kernel = world > 1 and args.stats_collective == "fused"
    stats_every = 1 if in_kernel else max(1, args.stats_every)
    if world > 1 and args.stats_overlap and args.stats_collective == "nccl":
        _lib.reserve_cta_slots(device, args.reserve_slots)
    wl.bind(StatsWindow(device, reducer, max(1, args.stats_every), overlap=args.stats_overlap))
    fused_pub, fused = None, None
    if in_kernel:
        # the headline at N > 1: north_star's "all-reduce the per-step episode statistics" taken literally -- EVERY step,
        # inside the control kernel: one extra CTA of each launch (the publisher) clears the previous step's accumulator,
        # stores its vector into every rank's mailbox over NVLink and sums the rows of the step before, while the other
        # CTAs stream the law (compute + collective in one launch; nothing at the tail of the kernel)
        fused_pub = PeerStatsReducer(device, lagged=True)
        acc2, reduced2 = [_lib.stats_buffer(device), _lib.stats_buffer(device)], _lib.stats_buffer(device)
        assert wl.sets % 2 == 0          # consecutive steps alternate the two accumulators
        fused = [wl.ctl.bind(wl.state[k], wl.tgt[k], wl.out[k], stats=acc2[k & 1], stats_prev=acc2[(k & 1) ^ 1],
                             publish=fused_pub, reduced=reduced2) for k in range(wl.sets)]
    loop = StepLoop(device, wl, fused=fused if in_kernel else None, use_graph=not args.no_graph)

    with ClockSampler(local_rank) as clocks:
        # sustained warm-up so the clock samples describe the loaded state of this very kernel
        # (every rank must run the SAME number of steps: the statistics all-reduce fires every k-th step, and ranks
        # whose clocks disagree on when the warm-up ends would enqueue different numbers of collectives and deadlock;
        # the continue / stop decision is therefore itself reduced over the ranks)
        t_end = time.perf_counter() + args.sustain_s
        while True:
            go = torch.tensor([1 if time.perf_counter() < t_end else 0], device=device, dtype=torch.int32)
            if world > 1:
                dist.all_reduce(go, op=dist.ReduceOp.MIN)
            if int(go.item()) == 0:
                break
            loop.run(208)                 # a multiple of the statistics window: every rank ends on a window boundary
            torch.cuda.synchronize(device)
        loop.run(args.warmup)
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(device)
        # Lead-in: `sets` untimed steps issued AFTER the synchronize, so the start event is recorded on a busy stream and
        # the host has enqueued the first timed replay before the device gets there.  Without it the region opens on an
        # idle device and counts the host's graph-launch latency (~25 us: 33.2 vs 32.0 us/step at 20 steps,
        # gpurun_out/t13) -- a cost of the bracket, not of a step.  The timed region is still EXACTLY `steps` steps
        # between two events on the launching stream; `cold_start_ms_per_step` below is the same region without lead-in.
        lead_in = 0 if args.no_lead_in else wl.sets
        loop.run(lead_in)
        launches0, replays0, gsteps0 = _lib.launch_count(), loop.replays, loop.graph_steps
        start.record()
        loop.run(args.steps)
        end.record()
        torch.cuda.synchronize(device)
        # kernels launched inside the timed region: C-ABI calls issued one by one + the kernels of the replayed graphs
        timed_replays, timed_gsteps = loop.replays - replays0, loop.graph_steps - gsteps0
        launches = _lib.launch_count() - launches0 + timed_gsteps
        # the same K steps opened on an idle device (synchronize immediately before the start event)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        c0.record()
        loop.run(args.steps)
        c1.record()
        torch.cuda.synchronize(device)
        cold = torch.tensor([c0.elapsed_time(c1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(cold, op=dist.ReduceOp.MAX)
        cold_ms_per_step = cold.item() / args.steps
        wl.window.finish()
        stats_check = None
        if in_kernel:      # the exchanged vector of a full step: every rank must read the global env count
            stats_check = {"reduced_n_env": float(reduced2[0].item()), "expected": float(total_envs), "timeouts": fused_pub.timeouts()}
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
        link_s = host_link_ceiling(device, hs, ht, hout, world)

        ms_per_step = ms_total / args.steps
        per_step_stats = None
        if world > 1 and not args.no_strong and collective_note is None:
            # the other forms of the exchange on the same loop (same 4-step graph replays), for comparison with the headline's
            def timed_loop(lp):
                lp.run(64)
                torch.cuda.synchronize(device)
                dist.barrier()
                s1, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s1.record()
                lp.run(args.steps)
                e1.record()
                torch.cuda.synchronize(device)
                t1 = torch.tensor([s1.elapsed_time(e1)], device=device, dtype=torch.float64)
                dist.all_reduce(t1, op=dist.ReduceOp.MAX)
                return t1.item() * 1e3 / args.steps

            forms = {}
            for name, every, lagged in (("peer_kernel_side_stream_every_16", 16, False), ("peer_kernel_side_stream_lagged_every_1", 1, True)):
                red_ = PeerStatsReducer(device, lagged=lagged)
                wl.bind(StatsWindow(device, red_, every, overlap=True))
                forms[name] = timed_loop(StepLoop(device, wl, use_graph=not args.no_graph))
                wl.window.finish()
            pub_ = PeerStatsReducer(device, lagged=True)      # measured again here, next to the others (the headline ran minutes earlier)
            acc1, reduced1 = [_lib.stats_buffer(device), _lib.stats_buffer(device)], _lib.stats_buffer(device)
            fz = [wl.ctl.bind(wl.state[k], wl.tgt[k], wl.out[k], stats=acc1[k & 1], stats_prev=acc1[(k & 1) ^ 1], publish=pub_,
                              reduced=reduced1) for k in range(wl.sets)]
            forms["in_kernel_publisher_cta_every_1"] = timed_loop(StepLoop(device, wl, fused=fz, use_graph=not args.no_graph))
            wl.bind(StatsWindow(device, None, 16))
            forms["no_exchange"] = timed_loop(StepLoop(device, wl, use_graph=not args.no_graph))
            per_step_stats = {"us_per_step_by_exchange_form": forms,
                              "note": "same PD loop, 1,048,576 envs per GPU; in_kernel_publisher_cta = b200ctl_pd_torque_published "
                                      "(statistics of step s published from one extra CTA of kernel s + 1 over NVLink, global sum readable "
                                      "two steps later); peer_kernel = the stand-alone b200ctl all-reduce kernel; NCCL in order measured "
                                      "33 us (every 16) / 67 us (every step) at N=2, profiles/r02_ab_stats_allreduce_n2.txt"}
        strong = None
        if world > 1 and not args.no_strong:
            _lib.reserve_cta_slots(device, args.reserve_slots if (args.stats_overlap and args.stats_collective == "nccl") else 0)
            strong = strong_scaling(device, rank, world, reducer, max(1, args.stats_every), args.stats_overlap, args.steps, ms_per_step,
                                    fused_pub=PeerStatsReducer(device, lagged=True) if in_kernel else None)

    value = total_envs * args.steps / (ms_total * 1e-3)
    per_gpu_rate = (hi - lo) / (ms_per_step * 1e-3)
    achieved = per_gpu_rate * PD_BYTES_PER_ENV / 1e9
    e2e_value = total_envs * e2e_steps / e2e_s.item()
    step_bytes = (hs.numel() + ht.numel() + hout.numel()) * 4

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": make_config(world, hi - lo, stats_every, stats_mode_name(args), wl.sets),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "pd_torque_vec4_kernel", "bytes_per_launch": (hi - lo) * PD_BYTES_PER_ENV,
                     "peak_source": peak_src, "of": "measured" if peak_src.startswith("measured") else "fallback"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": (hs.numel() + ht.numel()) * 4 * world,
                "d2h_bytes_per_step": hout.numel() * 4 * world, "steps": e2e_steps,
                "api": "test_isaacgym_b200.pd_control.pd_torque(host tensors) -> b200ctl_pd_torque_host",
                "checksum": e2e_checksum,
                "host_link_ceiling": {"env_steps_per_s": total_envs / link_s, "ms_per_step": link_s * 1e3,
                                      "gb_per_s_per_rank": step_bytes / link_s / 1e9, "gb_per_s_all_ranks": step_bytes * world / link_s / 1e9,
                                      "how": "the step's host-to-device and device-to-host bytes as plain async copies on two streams "
                                             "from the same pinned buffers, no kernel, all ranks at once (barrier-aligned), slowest rank"},
                "frac_of_host_link_ceiling": link_s / (e2e_s.item() / e2e_steps)},
        "gpu_launches": int(launches),
        "timed_region": {"lead_in_steps": lead_in, "cold_start_ms_per_step": cold_ms_per_step,
                         "how": "barrier + synchronize, `lead_in_steps` untimed steps, start event, exactly `steps` steps, end "
                                "event, synchronize + barrier; cold_start = the same without lead-in (idle device at the start "
                                "event: counts the host's graph-launch latency once)"},
        "step_issue": ("single bound calls (--no-graph)" if loop.graphs is None else
                       f"CUDA graphs of {loop.block} / {wl.sets} consecutive steps (rotating over the {wl.sets} buffer sets), replayed: "
                       f"{timed_replays} replays = {timed_gsteps} steps + "
                       f"{int(launches) - timed_gsteps} single launches in the timed region"),
        "clocks": clocks.summary(),
    }
    if strong is not None:
        line["strong"] = strong
    if per_step_stats is not None:
        line["per_step_stats"] = per_step_stats
    if stats_check is not None:
        line["stats_check"] = stats_check
    if collective_note is not None:
        line["stats_collective_note"] = collective_note
    traffic_file = os.path.join(ROOT, "profiles", "pd_traffic.json")
    if os.path.isfile(traffic_file):
        try:
            line["roofline"]["traffic"] = json.load(open(traffic_file))["dram_bytes_per_launch"]
        except Exception:
            pass
    if rank == 0 and world == 1 and not args.no_copy_probe:
        here = copy_bandwidth_here(device, local_rank)
        line["roofline"]["copy_here"] = here
        line["roofline"]["frac_of_sustained_copy_here"] = {"zeros": achieved / here["zeros"]["sustained_gbs"],
                                                            "random": achieved / here["random"]["sustained_gbs"]}
    if rank == 0 and world == 1 and not args.no_families:
        # give the headline's 1.6 GB of buffers back first: the family entries are then laid out in device memory as
        # in a process of their own (with the buffers alive the 262,144-env OSC entry read 53.9 us instead of 50.9)
        del loop, wl, hs, ht, hout, h
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
    ap.add_argument("--stats-overlap", action="store_true", help="all-reduce on a side stream (the default for the peer-memory collective)")
    ap.add_argument("--stats-in-order", action="store_true", help="all-reduce in order on the control stream (the default for NCCL)")
    ap.add_argument("--reserve-slots", type=int, default=1, help="CTA slots every persistent grid leaves free with --stats-overlap")
    ap.add_argument("--stats-collective", default="fused", choices=["fused", "peer", "peer-lagged", "nccl"],
                    help="N > 1: statistics exchanged every step inside the PD kernel (default), the library's stand-alone "
                         "all-reduce kernel over NVLink peer memory every --stats-every steps, its lagged form, or NCCL")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling measurement (N > 1)")
    ap.add_argument("--sustain-s", type=float, default=1.0, help="seconds of pre-load before the timed region (clock sampling)")
    ap.add_argument("--no-graph", action="store_true", help="issue every step as a single bound call instead of replaying 4-step CUDA graphs")
    ap.add_argument("--no-lead-in", action="store_true", help="open the timed region on an idle device (no untimed steps between the synchronize and the start event)")
    ap.add_argument("--no-copy-probe", action="store_true", help="skip the in-process copy-bandwidth probe (roofline.copy_here)")
    ap.add_argument("--no-families", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # NCCL's 640-thread CTA cannot sit next to the persistent control grids (it needs an empty SM): in order unless asked;
    # the library's own 64-thread all-reduce kernel can: side stream unless asked otherwise
    args.stats_overlap = args.stats_overlap or (args.stats_collective != "nccl" and not args.stats_in_order)
    return run_reference(args) if args.impl == "reference" else run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
