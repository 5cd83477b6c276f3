#!/usr/bin/env python
"""Summarise ncu outputs into small text files that can be committed under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv            > profiles/rNN_launches.md
  python tools/ncu_summary.py raw gpurun_out/prof.ncu-rep                 > profiles/rNN_kernels.md
  python tools/ncu_summary.py phases gpurun_out/prof.ncu-rep <kernel-regex> > profiles/rNN_phases.md
"""
import collections
import csv
import io
import re
import subprocess
import sys


def launches(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = collections.OrderedDict()
    for r in rows:
        agg.setdefault(r["Kernel Name"], []).append(float(r["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print(f"ncu --metrics gpu__time_duration.sum launch list: {len(rows)} launches, {tot / 1e3:.1f} us total "
          f"(cold-cache, serialised: compare shares)\n")
    print("| kernel | launches | avg us | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        name = re.sub(r"\(.*", "", k)[:90]
        print(f"| `{name}` | {len(v)} | {sum(v) / len(v) / 1e3:.2f} | {100 * sum(v) / tot:.1f}% |")


METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    ki = hdr.index("Kernel Name")
    print("ncu --set full --clock-control none, one row per captured launch\n")
    print("| kernel | " + " | ".join(f"{m} [{units[i]}]" for m, i in cols) + " |")
    print("|---|" + "---:|" * len(cols))
    for d in data:
        print(f"| `{re.sub(r'[(].*', '', d[ki])[:70]}` | " + " | ".join(d[i] for _, i in cols) + " |")


def phases(path, pattern):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "-k", f"regex:{pattern}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    tables, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "data": []}
            tables.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
            cur["data"].append(r)
    t = tables[0]
    hdr, data = t["hdr"], t["data"]
    si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[si] or 0) for r in data)
    print(f"{t['name']}\n{len(data)} SASS instructions, {tot} warp samples; segments end at cluster barriers (UCGABAR_WAIT)\n")
    print("| SASS range | samples | share | warp-instr executed | top stall reasons |\n|---|---:|---:|---:|---|")
    prev = 0
    marks = [i for i, r in enumerate(data) if "UCGABAR_WAIT" in r[src]] + [len(data) - 1]
    for m in marks:
        sub = data[prev:m + 1]
        n = sum(int(r[si] or 0) for r in sub)
        agg = sorted(((sum(int(r[i] or 0) for r in sub), hdr[i]) for i in stall), reverse=True)[:3]
        print(f"| {prev}-{m} | {n} | {100 * n / max(tot, 1):.1f}% | {sum(int(r[ie] or 0) for r in sub)} | "
              + ", ".join(f"{k[6:]} {v}" for v, k in agg if v) + " |")
        prev = m + 1
    agg = sorted(((sum(int(r[i] or 0) for r in data), hdr[i]) for i in stall), reverse=True)[:6]
    print("\nwhole kernel: " + ", ".join(f"{k[6:]} {v}" for v, k in agg))


if __name__ == "__main__":
    {"launches": launches, "raw": raw, "phases": phases}[sys.argv[1]](*sys.argv[2:])
