#!/usr/bin/env python
"""Per-SASS-line stall table of one kernel from an exported source page:
   ncu -i REP --page source --csv --kernel-name regex:K > src.csv ; python tools/ncu_lines.py src.csv [first last | top N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "Address"]
hdr = rows[hi[0]]
end = hi[1] - 1 if len(hi) > 1 else len(rows)
data = [r for r in rows[hi[0] + 1:end] if len(r) == len(hdr)]
ix = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
def line(k, r):
    st = sorted(((int(r[ix[n]] or 0), n[6:]) for n in stalls), reverse=True)
    st = " ".join(f"{n}:{v}" for v, n in st[:3] if v)
    return (f"{k:5d} smp {r[ix['# Samples']]:>5} exe {r[ix['Instructions Executed']]:>8} wf {r[ix['L1 Wavefronts Shared']] or '-':>7} "
            f"tag {r[ix['L1 Tag Requests Global']] or '-':>8} | {r[ix['Source']].strip()[:70]:70} | {st}")
if len(sys.argv) > 2 and sys.argv[2] == "top":
    n = int(sys.argv[3])
    for k, r in sorted(enumerate(data), key=lambda kr: -int(kr[1][ix['# Samples']]))[:n]:
        print(line(k, r))
    tot = {n[6:]: sum(int(r[ix[n]] or 0) for r in data) for n in stalls}
    print({k: v for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v})
    print("shared wavefronts", sum(int(r[ix['L1 Wavefronts Shared']] or 0) for r in data), "global tag requests",
          sum(int(r[ix['L1 Tag Requests Global']] or 0) for r in data))
else:
    a, b = int(sys.argv[2]), int(sys.argv[3])
    for k in range(a, b + 1):
        print(line(k, data[k]))
