#!/usr/bin/env python3
"""Turn the ncu artefacts brought back in gpurun_out/ into the small text summaries committed here.

    python profiles/summarise.py launches gpurun_out/launches_pd.csv > profiles/r01_launches_pd.txt
    python profiles/summarise.py full gpurun_out/prof_pd.ncu-rep    > profiles/r01_full_pd.txt

`launches`: per-kernel share of the `--metrics gpu__time_duration.sum` pass (cold-cache, serialised:
shares are meaningful, absolutes are not).  `full`: selected counters of an `ncu --set full` capture.
"""
import collections
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def launches(path):
    rows = [ln for ln in open(path) if ln.startswith('"')]
    agg = collections.defaultdict(list)
    for r in csv.DictReader(io.StringIO("".join(rows))):
        agg[r["Kernel Name"][:110]].append(float(r["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print(f"# {path}: {sum(len(v) for v in agg.values())} launches, total {tot / 1e3:.1f} us (ncu, serialised, cold cache)")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{sum(v) / tot * 100:6.2f}%  n={len(v):4d}  avg={sum(v) / len(v) / 1e3:9.2f} us  {k}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    for row in rows[2:]:
        print("--- kernel", row[head.index("Kernel Name")][:120])
        for w in WANT:
            if w in head:
                i = head.index(w)
                print(f"  {w:84s} {row[i]:>18s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
