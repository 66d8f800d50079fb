#!/usr/bin/env python
"""Prints the headline metrics + top stall reasons + hottest SASS lines of an .ncu-rep (reads with `ncu -i`)."""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.sum", "l1tex__data_pipe_lsu_wavefronts_mem_lg.sum", "lts__t_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum"]


def main(rep, ntop=22):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units, vals = rows[0], rows[1], rows[2]
    print("kernel:", vals[h.index("Kernel Name")][:100])
    for w in WANT:
        if w in h:
            i = h.index(w)
            print(f"  {w:70s} {vals[i]:>16s} {units[i]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]
    idx = {n: i for i, n in enumerate(h)}
    data = [r for r in rows[2:] if len(r) > idx["# Samples"] and r[idx["# Samples"]].isdigit()]
    tot = sum(int(r[idx["# Samples"]]) for r in data) or 1
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    agg = {s: sum(int(r[idx[s]]) for r in data if r[idx[s]].isdigit()) for s in stalls}
    print("  stalls:", ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:7]))
    print("  total instructions (warp-level):", sum(int(r[idx["Instructions Executed"]]) for r in data))
    for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:ntop]:
        print(f"   {100 * int(r[idx['# Samples']]) / tot:5.1f}% {r[idx['Instructions Executed']]:>9s}  {r[1][:95]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 22)
