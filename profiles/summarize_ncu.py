#!/usr/bin/env python
"""Summarise an ncu report (read here, no GPU needed) into profiles/: key raw metrics per
captured launch, the warp-stall histogram and the per-row SASS opcode mix of the kernel.

    python profiles/summarize_ncu.py gpurun_out/<name>.ncu-rep profiles/<out>.json [units_per_launch]
"""
import collections
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_active.min", "sm__cycles_active.avg", "sm__cycles_active.max", "sm__cycles_elapsed.max",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
]


def ncu(*args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True, check=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    units = float(sys.argv[3]) if len(sys.argv) > 3 else None
    rows = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "raw", "--csv"))))
    h = rows[0]
    col = {c: i for i, c in enumerate(h)}
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[col["Kernel Name"]]}
        for k in KEYS:
            if k in col:
                v = r[col[k]].replace(",", "")
                try:
                    d[k] = float(v)
                except ValueError:
                    d[k] = v
        launches.append(d)
    summary = {"report": rep, "units": {k: rows[1][col[k]] for k in KEYS if k in col}, "launches": launches}
    try:
        src = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "source", "--csv"))))
        hi = [i for i, r in enumerate(src) if r and r[0] == "Address"]
        hh = src[hi[0]]
        body = src[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else None)]
        c2 = {c: i for i, c in enumerate(hh)}
        stalls = [c for c in hh if c.startswith("stall_") and "Not Issued" not in c]
        tot, ops, samples, inst = collections.Counter(), collections.Counter(), 0, 0
        for r in body:
            if len(r) < len(hh):
                continue
            for s in stalls:
                tot[s] += int(r[c2[s]] or 0)
            n = int(r[c2["Instructions Executed"]] or 0)
            samples += int(r[c2["# Samples"]] or 0)
            inst += n
            t = r[c2["Source"]].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            ops[op] += n
        summary["first_launch_source_page"] = {
            "sass_instructions": len(body), "warp_instructions_executed": inst, "samples": samples,
            "stall_fraction": {s: round(v / max(samples, 1), 4) for s, v in tot.most_common() if v},
            "warp_instructions_per_unit": ({k: round(v / units, 2) for k, v in ops.most_common(30)} if units else None),
            "warp_instructions_per_unit_total": (round(inst / units, 2) if units else None),
        }
    except Exception as e:  # no source page in the report
        summary["first_launch_source_page"] = {"error": str(e)}
    with open(out, "w") as f:
        json.dump(summary, f, indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
