#!/usr/bin/env python
"""Summarise an Nsight Compute report (.ncu-rep) into the text kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/r01_masked_scan_general.txt

Prints, for the first profiled launch: duration, DRAM bytes, pipe utilisation, issue statistics, the warp
stall breakdown, and the sampled-stall distribution over code regions (regions = runs of SASS instructions
with the same execution count).  Needs `ncu` on PATH; works without a GPU.
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep = sys.argv[1]
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"# {rep}\n# kernel: {name}\n")
    col = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    for k in KEYS:
        if k in col:
            print(f"{k:86s} {col[k][1]:>16s} {col[k][0]}")
    print("\n# warp stall reasons (warps stalled per issued instruction)")
    for h, (u, v) in col.items():
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(v or 0) > 0.02:
            print(f"{h.split('issue_stalled_')[1].split('_per_issue')[0]:28s} {float(v):6.3f}")

    src = page(rep, "source")
    shdr, data = src[1], src[2:]
    ix = {h: i for i, h in enumerate(shdr)}
    stall_cols = [k for k in shdr if k.startswith("stall_") and "Not Issued" not in k]
    segs, cur = [], None
    for i, r in enumerate(data):
        ex = int(r[ix["Instructions Executed"]] or 0)
        if cur is None or abs(ex - cur["ex"]) > 0.02 * max(ex, cur["ex"], 1):
            cur = {"ex": ex, "start": i, "n": 0, "samples": 0, "st": collections.Counter(), "first": r[ix["Source"]][:44]}
            segs.append(cur)
        cur["n"] += 1
        cur["samples"] += int(r[ix["# Samples"]] or 0)
        for k in stall_cols:
            v = int(r[ix[k]] or 0)
            if v:
                cur["st"][k[6:]] += v
    tot = sum(s["samples"] for s in segs) or 1
    print(f"\n# sampled warp states by code region ({tot} samples; region = SASS run with one execution count)")
    for s in segs:
        if s["samples"] > 0.005 * tot:
            top = ", ".join(f"{k} {v}" for k, v in s["st"].most_common(4))
            print(f"sass[{s['start']:5d}+{s['n']:4d}] exec/inst {s['ex']:10d}  {100 * s['samples'] / tot:5.1f}%  {top}  | {s['first']}")


if __name__ == "__main__":
    main()
