#!/usr/bin/env python3
"""Executed-instruction mix of one kernel launch from an ncu report's SASS source page.

    python profiles/opmix.py gpurun_out/prof.ncu-rep <kernel-regex> [launch-index]
Prints warp-level executed instructions grouped by opcode (top 25) and the stall-sample share."""
import collections
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
lines = out.splitlines()
starts = [i for i, ln in enumerate(lines) if ln.startswith('"Kernel Name"')]
lines = lines[starts[0]:starts[1]] if len(starts) > 1 else lines[starts[0]:]
print(lines[0][:160])
rows = list(csv.DictReader(io.StringIO("\n".join(lines[1:]))))
ops, samples = collections.Counter(), collections.Counter()
for r in rows:
    if not r.get("Instructions Executed") or not r.get("Source"):
        continue
    src = r["Source"].strip()
    if src.startswith("@"):
        src = src.split(None, 1)[1]
    op = src.split()[0].split(".")[0]
    ops[op] += int(r["Instructions Executed"])
    samples[op] += int(r["# Samples"])
tot, stot = sum(ops.values()), max(1, sum(samples.values()))
print(f"total warp instructions executed: {tot}")
for op, n in ops.most_common(25):
    print(f"  {op:10s} {n:12d} {n / tot * 100:6.2f}%   stall samples {samples[op] / stot * 100:6.2f}%")
