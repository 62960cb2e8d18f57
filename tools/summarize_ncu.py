#!/usr/bin/env python3
"""Summarise an .ncu-rep capture (ncu -i ... --page raw/source --csv) into a small text file for profiles/."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes.sum.per_second", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def multi(rep, out):
    """A capture with many kernels: one metric block per distinct kernel name (its slowest launch)."""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn, dur = hdr.index("Kernel Name"), hdr.index("gpu__time_duration.sum")
    best = {}
    for r in data:
        if len(r) <= dur:
            continue
        name = r[kn]
        if name not in best or float(r[dur] or 0) > float(best[name][dur] or 0):
            best[name] = r
    lines = [f"# ncu --set full, one block per kernel (slowest launch of each) of {rep}", ""]
    for name, vals in sorted(best.items()):
        lines.append(f"kernel: {name}")
        grid_i, block_i = hdr.index("Grid Size") if "Grid Size" in hdr else None, hdr.index("Block Size") if "Block Size" in hdr else None
        if grid_i is not None:
            lines.append(f"  grid {vals[grid_i]}  block {vals[block_i]}")
        for i, h in enumerate(hdr):
            if h in KEYS:
                lines.append(f"  {h:73s} {vals[i]:>18s} {units[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(vals[i] or 0) >= 0.3:
                stalls.append((float(vals[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        lines.append("  stalls: " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:6]))
        lines.append("")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


def main():
    if sys.argv[1] == "--multi":
        return multi(sys.argv[2], sys.argv[3])
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    lines = [f"# ncu --set full summary of {rep}", f"kernel: {vals[hdr.index('Kernel Name')]}", ""]
    for i, h in enumerate(hdr):
        if h in KEYS:
            lines.append(f"{h:75s} {vals[i]:>18s} {units[i]}")
    lines.append("")
    lines.append("warp stall reasons (average warps stalled per issue-active cycle):")
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(vals[i] or 0) >= 0.1:
            lines.append(f"  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {float(vals[i]):6.2f}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr2, data = rows[hi], rows[hi + 1:]
    iex, ismp = hdr2.index("Instructions Executed"), hdr2.index("Warp Stall Sampling (All Samples)")

    def f(x):
        try:
            return float(x)
        except ValueError:
            return 0.0
    cand = [r for r in data if len(r) > iex and r[0].strip().isdigit()]
    cand.sort(key=lambda r: -f(r[ismp]))
    lines.append("")
    lines.append("hottest CUDA source lines (line | executed warp instructions | stall samples):")
    for r in cand[:25]:
        lines.append(f"  {r[0].strip():>5s} | {int(f(r[iex])):>11d} | {int(f(r[ismp])):>6d} | {r[1].strip()[:100]}")
    lines.append("")
    lines.append("CUDA source lines by executed warp instructions (line | executed | share):")
    total = sum(f(r[iex]) for r in cand) or 1.0
    for r in sorted(cand, key=lambda r: -f(r[iex]))[:40]:
        lines.append(f"  {r[0].strip():>5s} | {int(f(r[iex])):>11d} | {100.0 * f(r[iex]) / total:5.1f} % | {r[1].strip()[:90]}")
    # pipe utilisation (which issue port is the busy one) and the SASS opcode mix behind it
    lines.append("")
    lines.append("pipe utilisation:")
    for i, h in enumerate(hdr):
        if "pipe" in h and ("pct_of_peak_sustained_active" in h) and f(vals[i]) >= 1.0:
            lines.append(f"  {h:78s} {f(vals[i]):6.1f} {units[i]}")
    sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                          capture_output=True, text=True).stdout
    rows = list(csv.reader(sass.splitlines()))
    try:
        hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
        hdr3, data3 = rows[hi], rows[hi + 1:]
        iex3 = hdr3.index("Instructions Executed")
        isrc = hdr3.index("Source") if "Source" in hdr3 else 1
        ops = {}
        for r in data3:
            if len(r) <= iex3:
                continue
            toks = r[isrc].split()
            if not toks:
                continue
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            ops[op.rstrip(";")] = ops.get(op.rstrip(";"), 0.0) + f(r[iex3])
        tot = sum(ops.values()) or 1.0
        lines.append("")
        lines.append(f"SASS opcode mix by executed warp instructions (total {int(tot)}):")
        for op, n in sorted(ops.items(), key=lambda kv: -kv[1])[:45]:
            lines.append(f"  {op:28s} {int(n):>12d} {100.0 * n / tot:5.1f} %")
    except StopIteration:
        lines.append("(no SASS page)")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:30]))


if __name__ == "__main__":
    main()
