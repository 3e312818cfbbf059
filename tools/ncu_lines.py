#!/usr/bin/env python3
"""Per-source-line totals (stall samples, instructions executed) of one kernel from an ncu report,
using nvdisasm line info of the built library.
   python tools/ncu_lines.py rep.ncu-rep <mangled-kernel-substring> [top]"""
import csv, subprocess, sys, io, collections, os, re, tempfile, glob
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "eicos_b200", "libeicos_b200.so")], cwd=tmp, capture_output=True)
cub = [f for f in glob.glob(tmp + "/*.cubin") if os.path.basename(f).startswith("engine.")][0]
dis = subprocess.run(["nvdisasm", "-g", cub], capture_output=True, text=True).stdout.splitlines()
lines = []  # per instruction: (file, line)
inside = False; cur = ("?", 0)
for ln in dis:
    if ln.startswith("\t.section\t.text."):
        inside = kern in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
print("sass instructions: ncu", len(data), "nvdisasm", len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0])
for i, r in enumerate(data):
    key = lines[i] if i < len(lines) else ("?", 0)
    agg[key][0] += int(r[col["# Samples"]] or 0)
    agg[key][1] += int(r[col["Instructions Executed"]] or 0)
    agg[key][2] += 1
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
src = {}
def text(f, l):
    if f not in src:
        p = os.path.join(ROOT, "eicos_b200", "csrc", f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    return src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ""
print(f"total samples {ts} inst executed {ti}")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{key[0]:18s}:{key[1]:5d} smp {100*v[0]/ts:5.1f}%  inst {100*v[1]/ti:5.1f}%  sass {v[2]:4d}  {text(*key)}")
