#!/usr/bin/env python3
"""Turn the raw ncu outputs of tools/gpu_round.sh (gpurun_out/<tag>/) into the small text/JSON
summaries that are committed under profiles/.

    python tools/ncu_summary.py gpurun_out/r01a profiles/r01a
"""
import collections
import csv
import json
import os
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui].strip().replace("usecond", "us").replace("nsecond", "ns").replace("msecond", "ms"), 1e-6)
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    tot = sum(v[1] for v in agg.values())
    return [{"kernel": k, "launches": v[0], "total_ms": round(v[1], 3), "avg_ms": round(v[1] / v[0], 4),
             "share": round(v[1] / tot, 4)} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0]}
        for w in WANT:
            if w in hdr:
                d[w] = f"{r[hdr.index(w)]} {units[hdr.index(w)]}".strip()
        res.append(d)
    return res


def main():
    src, dst = sys.argv[1], sys.argv[2]
    os.makedirs(os.path.dirname(dst) or ".", exist_ok=True)
    summary = {"source": src}
    lp = os.path.join(src, "launches.csv")
    if os.path.exists(lp):
        summary["launch_list"] = {"command": "ncu --metrics gpu__time_duration.sum --clock-control none -k regex:eicos_ "
                                             "python bench.py --batch 4096 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline",
                                  "note": "cold-cache, serialised launches: compare shares, not absolutes",
                                  "kernels": launches(lp)}
    for name in ("prof_solve_kkt", "prof_ldl_factor"):
        rep = os.path.join(src, name + ".ncu-rep")
        if os.path.exists(rep):
            summary[name] = {"command": "ncu --set full --clock-control none --import-source on (same bench command, batch 4096)",
                             "launches": raw_page(rep)}
    for f in ("bench.json", "bench_reference.json"):
        p = os.path.join(src, f)
        if os.path.exists(p) and os.path.getsize(p):
            try:
                summary[f] = json.loads(open(p).read().strip().splitlines()[-1])
            except ValueError:
                pass
    json.dump(summary, open(dst + "_summary.json", "w"), indent=1)
    print(json.dumps(summary.get("launch_list", {}), indent=1)[:3000])
    for name in ("prof_solve_kkt", "prof_ldl_factor"):
        for l in summary.get(name, {}).get("launches", [])[:1]:
            print(name, json.dumps(l, indent=1))


if __name__ == "__main__":
    main()
