#!/usr/bin/env python
"""Isolated throughput of the fixpoint kernel (propagate_kernel): every resident block repeatedly
propagates a copy of the root store to its fixpoint.  This is the kernel the shared-memory roofline
of SURVEY.md 8(d) is stated for; bench.py embeds the same measurement as `fixpoint_kernel`.

  python tools/fixpoint_bench.py [--workload trains15] [--repeat 20] [--fp wac1] [--mem auto]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from turbo_b200 import abi  # noqa: E402

MEM = {"auto": abi.MEM_AUTO, "global": abi.MEM_GLOBAL, "store_shared": abi.MEM_STORE_SHARED, "tcn_shared": abi.MEM_TCN_SHARED,
       "store_cluster": abi.MEM_STORE_CLUSTER}


def measure(pb, repeat=20, fp="wac1", mem="auto", tpb=0, blocks=0, device=0, rounds=3, sm_mhz=None):
    from turbo_b200 import engine
    opts = dict(device=device, propagate_repeat=repeat, fixpoint=abi.FP_KINDS[fp], mem_kind=MEM[mem])
    if tpb:
        opts["threads_per_block"] = tpb
    if blocks:
        opts["or_blocks"] = blocks
    with engine.Solver(pb, **opts) as s:
        cfg = s.config()
        nb = cfg["num_blocks"]
        lb = np.tile(pb.lb, (nb, 1))
        ub = np.tile(pb.ub, (nb, 1))
        best = None
        for _ in range(rounds + 1):          # first round is the warm-up
            r = s.propagate_batch(lb, ub)
            st = r["stats"]
            if best is None or st["kernel_ms"] < best["kernel_ms"]:
                best = st
    secs = best["kernel_ms"] / 1e3
    ded = best["num_deductions"]
    smem_bytes = 24.0 * ded + 4.0 * best["bounds_narrowed"]
    clk = sm_mhz or 1965.0
    sms = engine.device_info(device)["sm_count"]
    peak = 128 * sms * clk * 1e6 / 1e9
    return {"workload_vars": pb.nvars, "workload_props": pb.nprops, "fixpoint": fp, "memory_configuration": abi.MEM_NAMES.get(cfg["mem_kind"]),
            "blocks": nb, "threads_per_block": cfg["threads_per_block"], "repeat": repeat, "kernel_ms": best["kernel_ms"],
            "propagations": ded, "sweeps": best["fixpoint_iterations"], "bounds_narrowed": best["bounds_narrowed"],
            "propagations_per_sec": ded / secs, "smem_gbs": smem_bytes / secs / 1e9, "smem_peak_gbs": peak,
            "smem_frac": smem_bytes / secs / 1e9 / peak, "props_per_clk_per_sm": ded / secs / (clk * 1e6) / sms}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="trains15")
    ap.add_argument("--repeat", type=int, default=20)
    ap.add_argument("--fp", default="wac1")
    ap.add_argument("--mem", default="auto")
    ap.add_argument("--tpb", type=int, default=0)
    ap.add_argument("--blocks", type=int, default=0)
    ap.add_argument("--rounds", type=int, default=3)
    a = ap.parse_args()
    from bench import load_workload
    pb, _ = load_workload(a.workload)
    print(json.dumps(measure(pb, a.repeat, a.fp, a.mem, a.tpb, a.blocks, rounds=a.rounds)))


if __name__ == "__main__":
    main()
