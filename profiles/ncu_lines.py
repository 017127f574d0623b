#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda` dump per source line:
instructions executed, stall samples and top stall reasons.  Usage: ncu_lines.py report.ncu-rep [topN]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
def fix(r):
    # source text with embedded quotes (inline asm) is split into extra cells: merge them back into the Source column
    extra = len(r) - len(hdr)
    return r if extra <= 0 else [r[0], ",".join(r[1:2 + extra])] + r[2 + extra:]
rows = rows[:hi + 1] + [fix(r) for r in rows[hi + 1:]]
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[0].isdigit():
        continue
    num = lambda v: int(float(v)) if v not in ("", "-") else 0
    inst = num(r[col["Instructions Executed"]])
    samp = num(r[col["# Samples"]])
    st = sorted(((num(r[col[s]]), s) for s in stall_cols), reverse=True)[:3]
    data.append((int(r[0]), r[1].strip()[:90], inst, samp, st))
ti = sum(d[2] for d in data); ts = sum(d[3] for d in data)
print(f"total warp-instructions {ti}, samples {ts}")
print("== by instructions")
for d in sorted(data, key=lambda d: -d[2])[:top]:
    print(f"{d[0]:5d} {100*d[2]/ti:5.1f}% inst {100*d[3]/max(ts,1):5.1f}% samp  {d[1]}   {[(s,n) for n,s in d[4] if n]}")
print("== by samples")
for d in sorted(data, key=lambda d: -d[3])[:top]:
    print(f"{d[0]:5d} {100*d[2]/ti:5.1f}% inst {100*d[3]/max(ts,1):5.1f}% samp  {d[1]}   {[(s,n) for n,s in d[4] if n]}")
if len(sys.argv) > 3:
    # phase aggregation: pass "name:lo-hi,name:lo-hi" (line ranges of the main source file)
    print("== by phase (samples%, inst%)")
    for spec in sys.argv[3].split(","):
        name, rng = spec.split(":"); lo, hi_ = map(int, rng.split("-"))
        sel = [d for d in data if lo <= d[0] <= hi_]
        print(f"{name:12s} samples {100*sum(d[3] for d in sel)/max(ts,1):5.1f}%  inst {100*sum(d[2] for d in sel)/ti:5.1f}%")
tot = {}
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[0].isdigit():
        continue
    for s in stall_cols:
        v = r[col[s]]
        tot[s] = tot.get(s, 0) + (int(float(v)) if v not in ("", "-") else 0)
print("== stall reasons overall:", sorted(((v, k) for k, v in tot.items() if v), reverse=True)[:8])
