#!/usr/bin/env python3
"""Condense an .ncu-rep (one kernel capture, --set full) into a small JSON/text summary for profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r01_x.json [--lines 25]
"""
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    res = {"report": rep.split("/")[-1], "kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            try:
                res[k] = {"value": float(vals[i]), "unit": units[i]}
            except ValueError:
                res[k] = {"value": vals[i], "unit": units[i]}
    sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(sass.splitlines()))
    if len(srows) > 2:
        h = srows[1]
        isrc, iex = h.index("Source"), h.index("Instructions Executed")
        ops = {}
        for r in srows[2:]:
            if len(r) <= iex:
                continue
            toks = r[isrc].split()
            if not toks:
                continue
            op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
            ops[op] = ops.get(op, 0) + int(r[iex] or 0)
        total = sum(ops.values()) or 1
        res["sass_opcode_mix_pct"] = {k: round(100.0 * v / total, 2) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:18]}
        res["tma_sass_executed"] = {k: ops.get(k, 0) for k in ("UTMALDG", "UBLKCP", "UTMASTG", "SYNCS")}
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
        f.write("\n")
    print(json.dumps(res, indent=1)[:1500])


if __name__ == "__main__":
    main()
