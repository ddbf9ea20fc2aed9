#!/usr/bin/env python
"""Curated text summary of one kernel of an .ncu-rep file (what profiles/*.txt are made from).

    python tools/ncu_summary.py REPORT.ncu-rep "title line" > profiles/rNN_xxx.txt

Reads `ncu -i REPORT --page raw --csv` and, when the report holds source counters,
`--page source --csv` (per-SASS-instruction samples, aggregated by opcode)."""
import collections
import csv
import io
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    print(title)
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    seen = {}
    for r in rows[2:]:
        kern = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        grid = r[hdr.index("launch__grid_size")] if "launch__grid_size" in hdr else ""
        if (kern, grid) in seen:            # repeated launches of the same kernel and grid: the first one stands
            seen[(kern, grid)] += 1
            continue
        seen[(kern, grid)] = 1
        print(f"\n== {kern}")
        for i, h in enumerate(hdr):
            stall = "issue_stalled" in h and h.endswith("per_issue_active.ratio")
            if h in KEEP or (stall and float(r[i] or 0) >= 0.03):
                print(f"  {h:88s} {units[i]:14s} {r[i]}")
    src = page(rep, "source")
    if len(src) > 3 and "Instructions Executed" in src[1]:
        h = {k: i for i, k in enumerate(src[1])}
        by = collections.defaultdict(lambda: [0, 0, 0])
        for r in src[2:]:
            if len(r) <= max(h["Instructions Executed"], h["# Samples"], h["L1 Wavefronts Shared"]) or not r[h["Instructions Executed"]].strip().isdigit():
                continue                    # (header lines of further kernels in a multi-kernel report)
            m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[1])
            op = m.group(1) if m else "?"
            if op in ("LDS", "STS", "LDG", "STG"):
                op += ".128" if ".128" in r[1] else (".64" if ".64" in r[1] else "")
            by[op][0] += int(r[h["Instructions Executed"]])
            by[op][1] += int(r[h["# Samples"]])
            by[op][2] += int(r[h["L1 Wavefronts Shared"]] or 0)
        tex = sum(v[0] for v in by.values())
        tsm = sum(v[1] for v in by.values())
        print("\n  executed warp instructions by opcode (source counters of the same capture, all profiled launches)")
        print("  opcode      %instructions  %stall-samples  smem wavefronts/instruction")
        for k, v in sorted(by.items(), key=lambda kv: -kv[1][0])[:14]:
            print(f"  {k:10s} {100 * v[0] / tex:12.1f} {100 * v[1] / max(1, tsm):15.1f} {v[2] / max(1, v[0]):12.2f}")


if __name__ == "__main__":
    main()
