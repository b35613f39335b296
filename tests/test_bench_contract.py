"""bench.py contract checks that run without a GPU: the reference arm (CPU oracle) prints one JSON line with the
keys the driver reads; the GPU arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600)


def test_reference_arm_json_line():
    r = _run("--impl", "reference", "--steps", "2", "--warmup", "3")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "env-steps/s" and line["steps"] == 2
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0
    assert "workload" in line["config"] and "model" not in line["config"]
    # the two arms describe ONE workload: the reference arm's config is built by the same function as the repo arm's
    import bench
    assert line["config"] == bench.make_config(1, bench.ENVS_PER_GPU, 16, "x")
    assert set(line["config"]) == {"workload", "envs_per_gpu", "num_dofs", "global_envs", "parallelism",
                                   "stats_allreduce_every", "stats_allreduce", "l2_policy"}


def test_rotation_follows_touched_bytes():
    """Buffer rotation is sized by the bytes a launch TOUCHES: twice the 126 MB L2 between two uses of a set."""
    import bench
    assert bench.sets_for(16_384 * 958) == 17            # OSC at C3: 15.7 MB touched of a 47 MB set
    assert bench.sets_for(16_384 * 460, hi=40) == 36     # IK at C3
    assert bench.sets_for(1_048_576 * 192) == 3          # PD headline: inputs alone exceed L2
    assert bench.sets_for(1, hi=40) == 40 and bench.sets_for(10 ** 12) == 3


def test_reference_arm_non_zero_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_gpu_arm_refuses_without_a_device():
    r = _run("--steps", "1")
    assert r.returncode != 0
    assert "no CPU path" in r.stdout


def test_stdout_carries_the_json_line_only():
    """Native libraries write to file descriptor 1 behind Python's back (NCCL's version banner under
    NCCL_DEBUG=VERSION preceded the JSON line of the N > 1 runs): after ``claim_stdout`` such output lands on stderr
    and the line alone on the original stdout."""
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.claim_stdout(); "
            "os.write(1, b'NCCL version 2.28.9+cuda12.9\\n'); print('python-level noise'); "
            "bench.emit({'metric': 'm', 'value': 1.0})" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.count("\n") == 1 and json.loads(r.stdout) == {"metric": "m", "value": 1.0}
    assert "NCCL version" in r.stderr and "python-level noise" in r.stderr


@pytest.mark.parametrize("window_form,every", [(False, 16), (True, 16), (True, 4), (True, 3)])
def test_step_loop_blocks_issue_the_same_steps_as_single_calls(monkeypatch, window_form, every):
    """The headline loop replays 20-step and 4-step graphs and falls back to single bound calls: whatever mix it picks,
    the (buffer set, accumulator bank) sequence must be the one single calls would issue, the step counts must add up,
    and with a statistics exchange between steps (window forms) no replayed block may straddle a window boundary."""
    import bench
    import test_isaacgym_b200.graph as graph_mod

    log = []

    class FakeGraph:
        def __init__(self, calls, device, warmup=0, restore=()):
            self.calls = calls

        def __call__(self):
            FakeWindow.block_open = len(self.calls)
            for c in self.calls:
                c()

    class FakeWindow:
        block_open = 0      # steps of the replay in flight whose step_done() has not been called yet

        def __init__(self):
            self.reducer, self.every, self.cur, self._steps, self.boundaries_inside_blocks = (object() if window_form else None), every, 0, 0, 0

        def step_done(self):
            self._steps += 1
            if FakeWindow.block_open:
                FakeWindow.block_open -= 1
            if self._steps % self.every == 0:
                if FakeWindow.block_open:       # an exchange would be due while later steps of the block are already enqueued
                    self.boundaries_inside_blocks += 1
                self.cur ^= 1

    class FakeWorkload:
        sets = 4

        def __init__(self):
            self.window = FakeWindow()
            self.calls = [[(lambda k=k, b=b: log.append((k, b))) for b in (0, 1)] for k in range(self.sets)]

        def step(self, i):
            self.calls[i % self.sets][self.window.cur]()
            self.window.step_done()

    monkeypatch.setattr(graph_mod, "StepGraph", FakeGraph)
    for ks in ([20], [3, 20, 23, 1, 7, 44], [400], [5, 5, 5, 64]):
        log.clear()
        wl = FakeWorkload()
        loop = bench.StepLoop(None, wl)
        assert loop.block == (20 if not window_form or every >= 20 else max(4, 4 * (min(20, every) // 4)) if every >= 4 else 4)
        issued = wl.sets
        for k in ks:
            before = len(log)
            loop.run(k)
            issued += k
            assert len(log) - before == k and loop.pos == issued
        assert [s for s, _ in log] == [i % wl.sets for i in range(issued)]
        if window_form:
            assert wl.window.boundaries_inside_blocks == 0
            assert [b for _, b in log] == [(i // every) & 1 for i in range(issued)]
        assert loop.graph_steps <= issued and (every < 4 and window_form) == (loop.graphs is None)
