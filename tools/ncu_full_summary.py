#!/usr/bin/env python3
"""Summary of `ncu --set full` captures (one launch each) for profiles/: duration, occupancy limits, issue activity,
stall reasons per issued instruction, DRAM bytes, shared-memory wavefronts, and the hottest source lines.

    python tools/ncu_full_summary.py gpurun_out/r02b profiles/r02b_ncu_full.json eicos_solve_kkt eicos_ldl_factor ...
"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "not_selected", "math_pipe_throttle", "mio_throttle", "lg_throttle",
          "branch_resolving", "no_instruction", "barrier", "membar", "dispatch_stall"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def main():
    src, dst, kernels = sys.argv[1], sys.argv[2], sys.argv[3:]
    res = {"source": src, "command": "ncu --set full --clock-control none --import-source on -k regex:^<kernel>$ -s 3 -c 1 "
                                     "python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline   (batch 65536 = the bench batch)"}
    for k in kernels:
        rep = f"{src}/prof_{k}.ncu-rep"
        d = raw(rep)
        rec = {key: " ".join(d[key]) for key in KEYS if key in d}
        rec["stalls_per_issue"] = {s: d[f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"][0]
                                   for s in STALLS if f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio" in d}
        lines = subprocess.run([sys.executable, "tools/ncu_lines.py", rep, k + "E", "10"], capture_output=True, text=True).stdout
        rec["hot_source_lines"] = [l for l in lines.splitlines()[2:]]
        res[k] = rec
    json.dump(res, open(dst, "w"), indent=1)
    print(json.dumps(res, indent=1)[:3000])


if __name__ == "__main__":
    main()
