#!/usr/bin/env python
"""Decode the scheduling control fields of a cuobjdump -sass listing (stall count, write / read scoreboard, wait mask):
   cuobjdump -sass -fun NAME obj.o > f.txt ; python tools/sass_ctrl.py f.txt FIRST_HEX LAST_HEX [grep]"""
import re, sys
lines = open(sys.argv[1]).read().split('\n')
a, b = int(sys.argv[2], 16), int(sys.argv[3], 16)
pat = sys.argv[4] if len(sys.argv) > 4 else None
i = 0
while i < len(lines):
    m = re.match(r'\s+/\*([0-9a-f]{4})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/', lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r'\s+/\* (0x[0-9a-f]+) \*/', lines[i + 1])
        if m2:
            hi = int(m2.group(1), 16)
            stall, wr, rd, wait = (hi >> 41) & 0xf, (hi >> 46) & 7, (hi >> 49) & 7, (hi >> 52) & 0x3f
            addr = int(m.group(1), 16)
            if a <= addr <= b and (pat is None or re.search(pat, m.group(2)) or wait):
                w = ",".join(str(k) for k in range(6) if wait >> k & 1)
                print(f"{addr:04x} st{stall:2d} wr{wr if wr != 7 else '-'} rd{rd if rd != 7 else '-'} wait[{w:5}] | {m.group(2).strip()[:100]}")
            i += 2
            continue
    i += 1
