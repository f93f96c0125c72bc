#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: top SASS lines by stall samples + stall mix."""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; body = []
for r in rows[hi + 1:]:
    if r and r[0] == "Kernel Name":
        break          # first kernel of the export only
    if len(r) == len(hdr) and r[0].startswith("0x"):
        body.append(r)
si = hdr.index("# Samples"); src = hdr.index("Source"); ex = hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si] or 0) for r in body)
print(f"{len(body)} SASS lines, {tot} samples")
mix = {hdr[i]: sum(int(r[i] or 0) for r in body) for i in stalls}
print("stall mix:", {k: v for k, v in sorted(mix.items(), key=lambda kv: -kv[1]) if v})
for n, r in sorted(enumerate(body), key=lambda nr: -int(nr[1][si] or 0))[:top]:
    why = {hdr[i][6:]: int(r[i]) for i in stalls if r[i] and int(r[i])}
    print(f"{n:5d} {int(r[si]):6d} x{r[ex]:>7}  {r[src].strip()[:70]:70s} {why}")
