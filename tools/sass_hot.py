#!/usr/bin/env python
"""Bucket the SASS-level sampling of one kernel of an .ncu-rep (ncu --page source --csv) into regions of 100
instructions and list the hottest instructions.  python tools/sass_hot.py REP KERNEL_REGEX [bucket]"""
import csv
import subprocess
import sys
from collections import Counter

rep, kre = sys.argv[1], sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 100
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "Address"]
h0 = hi[0]
hdr = rows[h0]
end = hi[1] - 1 if len(hi) > 1 else len(rows)
data = [r for r in rows[h0 + 1:end] if len(r) == len(hdr)]
iS, iN, iI = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot_s = sum(int(r[iN]) for r in data)
tot_i = sum(int(r[iI]) for r in data)
print(rows[h0 - 1][1] if h0 else "", "| samples", tot_s, "| warp-inst", tot_i, "| sass", len(data))
KEYS = ("LDTM", "UTCHMMA", "LDG.E.128", "BAR.SYNC", "TRYWAIT", "MUFU.EX2", "REDUX", "STS.128", "UTCBAR", "STG.E.128", "SHFL", "LDL", "STL",
        "HMMA", "NANOSLEEP", "FFMA", "ATOM")
for b in range(0, len(data), bucket):
    seg = data[b:b + bucket]
    ss, ii = sum(int(r[iN]) for r in seg), sum(int(r[iI]) for r in seg)
    if ss / max(tot_s, 1) > 0.004 or ii / max(tot_i, 1) > 0.004:
        c = Counter(k for r in seg for k in KEYS if k in r[iS])
        print(f"{b:5d} samples {100 * ss / tot_s:5.1f}%  inst {100 * ii / tot_i:5.1f}%  {dict(c)}")
for k, r in sorted(enumerate(data), key=lambda kr: -int(kr[1][iN]))[:18]:
    print(k, r[iN], r[iI], r[iS].strip()[:100])
