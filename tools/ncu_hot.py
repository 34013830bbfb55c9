#!/usr/bin/env python
"""Where a kernel's samples go, from an .ncu-rep (read here, no GPU needed): the SASS instructions with the most
stall samples, each with its dominant stall reasons and execution count.
  python tools/ncu_hot.py gpurun_out/prof.ncu-rep [N]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, data = rows[1], rows[2:]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[isamp]) for r in data)
    totex = sum(int(r[iex]) for r in data)
    print(f"samples {tot}  executed warp instructions {totex}")
    order = sorted(range(len(data)), key=lambda k: -int(data[k][isamp]))[:n]
    print("--- top instructions by stall samples (index = position in the kernel)")
    for k in sorted(order):
        r = data[k]
        reasons = sorted(((int(r[i]), hdr[i][6:]) for i in stall), reverse=True)[:2]
        print(f"{k:6d} {100 * int(r[isamp]) / tot:5.2f}%  ex {int(r[iex]):>10d}  {' '.join(f'{b}:{a}' for a, b in reasons if a):32s} {r[isrc].strip()[:70]}")


if __name__ == "__main__":
    main()
