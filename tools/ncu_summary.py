#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into the text committed under profiles/:
key raw metrics, executed-opcode histogram and warp-stall breakdown from the source page.
  python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<kernel>_ncu.txt
"""
import csv
import io
import subprocess
import sys
from collections import Counter

RAW = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
       "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
       "sm__warps_active.avg.pct_of_peak_sustained_active",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
       "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_atom.sum",
       "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
        for w in RAW:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:78s} {vals[i]:>18s} {units[i]}")
    rows = page(rep, "source")
    hdr, data = rows[1], rows[2:]
    ia, isrc, ith = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Thread Instructions Executed")
    tot = sum(int(r[ia]) for r in data)
    tth = sum(int(r[ith]) for r in data)
    print(f"\nexecuted warp instructions: {tot}   average active threads per instruction: {tth / max(tot, 1):.1f}")
    c = Counter()
    for r in data:
        t = r[isrc].strip().split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        c[op] += int(r[ia])
    print("opcode histogram (share of executed warp instructions):")
    for op, n in c.most_common(18):
        print(f"  {op:12s} {100 * n / tot:5.1f}%")
    cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tots = {hdr[i]: sum(int(r[i]) for r in data) for i in cols}
    alls = sum(tots.values())
    print("warp stall sampling (all samples):")
    for k, v in sorted(tots.items(), key=lambda kv: -kv[1])[:10]:
        print(f"  {k:26s} {100 * v / max(alls, 1):5.1f}%")


if __name__ == "__main__":
    main()
