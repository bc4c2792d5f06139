#!/usr/bin/env python
"""Hot instructions of an ncu report's source page (SASS view): samples, executed count, top stall reasons.
    python tools/ncu_hot.py report.ncu-rep [top_n] [kernel-id]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
agg = {}
for r in body:
    for s in stalls:
        agg[s] = agg.get(s, 0) + int(r[ix[s]] or 0)
print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:10])
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = body[i]
    st = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stalls), reverse=True)[:3]
    print(f"{i:5d} {r[ix['# Samples']]:>6s} exec={r[ix['Instructions Executed']]:>8s} {r[ix['Source']].strip()[:70]:70s} {st}")
