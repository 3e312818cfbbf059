#!/usr/bin/env python3
"""Hot spots of one kernel from an ncu report: stall-reason totals and the SASS instructions with
the most stall samples (with a few lines of context).   python tools/ncu_hot.py rep.ncu-rep [top]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
for r in data:
    for h in stall_cols:
        tot[h] += int(r[col[h]] or 0)
ns = sum(int(r[col["# Samples"]] or 0) for r in data)
print("total samples", ns, "instructions", len(data), "inst executed", sum(int(r[col["Instructions Executed"]] or 0) for r in data))
for h, v in tot.most_common(8):
    print(f"  {h:28s} {v:8d} {100.0 * v / max(ns, 1):5.1f}%")
idx = sorted(range(len(data)), key=lambda i: -int(data[i][col["# Samples"]] or 0))[:top]
for i in sorted(idx):
    r = data[i]
    st = {h[6:]: int(r[col[h]] or 0) for h in stall_cols if int(r[col[h]] or 0) > 0}
    stt = ", ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{i:6d} {int(r[col['# Samples']]):6d} exec={r[col['Instructions Executed']]:>8s}  {r[col['Source']].strip():60s} {stt}")
