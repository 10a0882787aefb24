#!/usr/bin/env python
"""Summarise an `ncu --set full --import-source on` report into a small text file for profiles/.

usage: tools/summarise_ncu.py gpurun_out/prof_X.ncu-rep profiles/X.txt [--top 40]

Writes (a) the headline raw metrics (duration, DRAM bytes, issue utilisation, stall mix, occupancy
limits, shared-memory bank conflicts) and (b) the source lines with the most warp-stall samples,
with their dominant stall reasons.  Runs here (no GPU needed)."""
import csv
import io
import re
import subprocess
import sys

RAW_KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "smsp__sass_average_branch_targets_threads_uniform.pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
]


def ncu(*args):
    return subprocess.run(["ncu", *args], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    lines = []
    raw = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "raw", "--csv"))))
    hdr, units = raw[0], raw[1]
    for k, row in enumerate(raw[2:]):
        d = dict(zip(hdr, row))
        lines.append(f"== launch {k}: {d.get('Kernel Name', '?')}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        u = dict(zip(hdr, units))
        for key in RAW_KEYS:
            if key in d and d[key] != "":
                lines.append(f"  {key:75s} {d[key]:>16s} {u[key]}")
        stalls = [(float(d[h]), h) for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and d[h]]
        lines.append("  warp stall mix (warps stalled per issue-active cycle):")
        for v, h in sorted(stalls, reverse=True)[:8]:
            lines.append(f"    {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:8.3f}")
    # source hot spots (CUDA lines aggregated by ncu)
    src = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"))))
    cur_file, cols, hot, total = None, None, [], 0
    for r in src:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1]
        elif r and r[0] == "Line No":
            cols = r
        elif cols and len(r) == len(cols) and r[0].isdigit():
            d = dict(zip(cols[4:], r[4:]))
            try:
                s = int(d.get("# Samples", "0") or 0)
            except ValueError:
                continue
            total += s
            if s:
                st = sorted(((int(v), k) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)),
                            reverse=True)[:3]
                hot.append((s, cur_file, int(r[0]), r[1].strip(), d.get("Instructions Executed", ""), st))
    hot.sort(reverse=True)
    lines.append("")
    lines.append(f"== source hot spots: top {top} lines by warp-stall samples (total samples {total})")
    for s, f, ln, text, ninst, st in hot[:top]:
        sm = ", ".join(f"{k[6:]} {v}" for v, k in st)
        lines.append(f"  {100.0 * s / max(total, 1):5.1f}%  {f.split('/')[-1]}:{ln:<4d} inst {ninst:>10s}  [{sm}]")
        lines.append(f"          {text[:150]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
