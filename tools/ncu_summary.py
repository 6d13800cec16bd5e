#!/usr/bin/env python
"""Summarise ncu outputs into small text files that can be committed under profiles/.

  python tools/ncu_summary.py launches <launches.csv>            -> per-kernel count / total / share table
  python tools/ncu_summary.py full <report.ncu-rep> [regex]      -> key metrics of every profiled launch
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct",
    # the FP64 tensor-core (DMMA) sub-pipe and the FP64 pipe as cycle counters (tools/gpu_profile_round.sh adds them to --set full)
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_fp64.sum",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        k = re.sub(r"\(.*", "", r[ki])[:70]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v for _, v in agg.values())
    print(f"{'kernel':72s} {'n':>5s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}")
    for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:72s} {n:5d} {v / 1e6:10.3f} {v / tot:7.3f} {v / n / 1e3:9.1f}")
    print(f"{'TOTAL':72s} {sum(n for n, _ in agg.values()):5d} {tot / 1e6:10.3f}")


def full(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if pattern and not re.search(pattern, name):
            continue
        print(f"== {name}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"   {k:72s} {r[i]:>18s} {units[i]}")


def edge(summary_path, edges_per_launch):
    """profiles/rNN_ncu_edge_update.json (what bench.py quotes under roofline.ncu / roofline.traffic) from a summary
    written by `full`: the FIRST launch of each of the three edge-update kernels."""
    import json

    want = {"lq_chol_warp_kernel": None, "svd_warp_kernel": None, "recompose_chol_kernel": None}
    cur = None
    for line in open(summary_path):
        if line.startswith("== "):
            cur = next((k for k in want if k in line and want[k] is None), None)
            if cur:
                want[cur] = {}
        elif cur and line.strip():
            parts = line.split()
            try:
                want[cur][parts[0]] = float(parts[1].replace(",", ""))
            except (ValueError, IndexError):
                pass
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}

    # bytes need their units: read them again with the unit column
    dram_bytes = {k: 0.0 for k in want}
    cur, seen = None, set()
    for line in open(summary_path):
        if line.startswith("== "):
            cur = next((k for k in want if k in line and k not in seen), None)
            if cur:
                seen.add(cur)
        elif cur and ("dram__bytes_read.sum" in line or "dram__bytes_write.sum" in line):
            parts = line.split()
            dram_bytes[cur] += float(parts[1].replace(",", "")) * scale.get(parts[2], 1.0)
    get = lambda key: {k: v.get(key) for k, v in want.items() if v}  # noqa: E731
    out = {
        "source": f"{summary_path} (ncu --set full + DMMA sub-pipe counters, bench.py --batch {edges_per_launch}, one launch of each "
                  "kernel of a 1-edge level; tools/gpu_profile_round.sh)",
        "edges_per_launch_triple": edges_per_launch,
        "dram_bytes": dram_bytes,
        "kernel_time_us": get("gpu__time_duration.sum"),
        "fp64_pipe_pct": get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "dmma_pipe_pct": get("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"),
        "dram_throughput_pct": get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "l2_hit_pct": get("lts__t_sector_hit_rate.pct"),
        "dram_bytes_per_edge_update": sum(dram_bytes.values()) / edges_per_launch,
        "algorithmic_bytes_per_edge_update": 262656,
    }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "edge":
        edge(sys.argv[2], int(sys.argv[3]))
    elif sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
