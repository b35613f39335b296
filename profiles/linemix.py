#!/usr/bin/env python3
"""Executed-instruction counts of one kernel launch aggregated by CUDA source line.

    python profiles/linemix.py <prof.ncu-rep> <kernel-regex> <launch-skip> <lib.so> [top]

ncu's CSV source page is SASS-only, so the per-instruction "Instructions Executed" column is joined with the
file/line markers of `nvdisasm -g` on the cubin extracted from the SAME build of the library (instruction order
is identical; the join is by instruction index within the function and is checked opcode by opcode).  Inlined
device functions are attributed to their own file:line (innermost frame); libdevice code has no line info and
is attributed to the last line seen before it ("<- libdevice" rows usually follow a call site).
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def ncu_rows(rep, rx, skip):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx,
                          "--launch-skip", str(skip), "--launch-count", "1"], capture_output=True, text=True).stdout
    lines = out.splitlines()
    starts = [i for i, ln in enumerate(lines) if ln.startswith('"Kernel Name"')]
    lines = lines[starts[0]:starts[1]] if len(starts) > 1 else lines[starts[0]:]
    name = next(csv.reader([lines[0]]))[1]
    rows = [r for r in csv.DictReader(io.StringIO("\n".join(lines[1:]))) if r.get("Instructions Executed") and r.get("Source")]
    return name, rows


def disasm_lines(lib, demangled_name):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    def norm(x):
        x = re.sub(r"\((int|bool|unsigned int|long)\)", "", x.replace("void ", ""))
        x = x.replace("true", "1").replace("false", "0")
        return re.sub(r"\s", "", x)

    # template arguments may contain parentheses ("<(int)0>"): cut the parameter list at the LAST "(" group start
    head = norm(demangled_name)
    want = head[:head.index(">(") + 1] if ">(" in head else head[:head.index("(")]
    base = re.sub(r"<.*", "", want).split("::")[-1]
    for cub in sorted(os.listdir(tmp)):
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
        funcs = re.split(r"^//-+ \.text\.(\S+) -+$", txt, flags=re.M)
        for i in range(1, len(funcs), 2):
            mangled, body = funcs[i], funcs[i + 1]
            if base not in mangled:
                continue
            dem = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
            if not norm(dem).startswith(want):
                continue
            loc, out = ("?", 0), []
            for ln in body.splitlines():
                m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
                if m:
                    loc = (os.path.basename(m.group(1)), int(m.group(2)))
                    continue
                m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
                if m:
                    out.append((int(m.group(1), 16), m.group(2).strip(), loc))
            return out
    raise SystemExit(f"no function matching {want!r} in {lib}")


def main():
    rep, rx, skip, lib = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
    name, rows = ncu_rows(rep, rx, skip)
    dis = disasm_lines(lib, name)
    if len(dis) != len(rows):
        print(f"# WARNING: {len(rows)} profiled vs {len(dis)} disassembled instructions: different builds?")
    by_line, mism = collections.Counter(), 0
    for r, (off, text, loc) in zip(rows, dis):
        op_p = r["Source"].strip().split()[0 if not r["Source"].strip().startswith("@") else 1].split(".")[0]
        op_d = text.split()[0 if not text.startswith("@") else 1].split(".")[0]
        mism += op_p != op_d
        by_line[loc] += int(r["Instructions Executed"])
    tot = sum(by_line.values())
    print(f"# {name[:120]}\n# {tot} warp instructions executed; opcode mismatches in the join: {mism}")
    srcs = {}
    for (f, ln), n in by_line.most_common(top):
        path = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", f)
        if f not in srcs and os.path.isfile(path):
            srcs[f] = open(path).read().splitlines()
        text = srcs.get(f, [""] * (ln + 1))[ln - 1].strip()[:90] if f in srcs and ln - 1 < len(srcs[f]) else ""
        print(f"{n / tot * 100:6.2f}%  {n:10d}  {f}:{ln:<4d} {text}")


if __name__ == "__main__":
    main()
