#!/usr/bin/env python
"""Warp instructions executed per CUDA source line of one kernel: joins `nvdisasm --print-line-info` of the cubin (order of
the instructions) with the per-instruction counts of an exported ncu source page.
   python tools/sass_by_line.py LINES.txt MANGLED_NAME SRC.csv [divisor]"""
import csv, re, sys
from collections import Counter
txt = open(sys.argv[1]).read().split('\n')
name = sys.argv[2]
div = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
start = next(i for i, l in enumerate(txt) if l.startswith(f".text.{name}:"))
cur = None
lines = []
for l in txt[start + 1:]:
    if l.startswith("//-----") or l.startswith(".text."):
        break
    m = re.match(r'\s*//## File "(.*?)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4}\*/', l):
        lines.append((cur, l.split('*/', 1)[1].strip()))
rows = list(csv.reader(open(sys.argv[3])))
hi = [i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "Address"]
hdr = rows[hi[0]]
data = [r for r in rows[hi[0] + 1:] if len(r) == len(hdr)]
iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
n = min(len(lines), len(data))
if len(lines) != len(data):
    print("warning: length mismatch", len(lines), len(data))
cnt, smp = Counter(), Counter()
for k in range(n):
    cnt[lines[k][0]] += int(data[k][iI])
    smp[lines[k][0]] += int(data[k][iS])
tot = sum(cnt.values())
print(f"total warp instructions {tot} ({tot / div:.1f} per unit), samples {sum(smp.values())}")
src_cache = {}
for (f, ln), c in sorted(cnt.items(), key=lambda kv: -kv[1])[:60]:
    try:
        if f not in src_cache:
            import glob
            cand = glob.glob(f"mingraph_unet_b200/csrc/{f}")
            src_cache[f] = open(cand[0]).read().split('\n') if cand else []
        text = src_cache[f][ln - 1].strip()[:90] if src_cache[f] else ""
    except Exception:
        text = ""
    print(f"{c / div:8.1f} inst {100 * smp[(f, ln)] / max(1, sum(smp.values())):5.1f}% smp  {f}:{ln}  {text}")
