#!/usr/bin/env python
"""Where the conv kernel's warps spend their samples (read here, no GPU): per launch of an `ncu --set full
--import-source on` report, the share of stall samples that falls into barrier / mbarrier waits (idle role warps),
into the epilogue's pass body (instructions executed once per warp pass, found through the I2FP count) and the
stall mix inside that body, plus the instructions one pass issues.
    python tools/ncu_roles.py report.ncu-rep [launch indices ...] > profiles/x_roles.txt"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
launches = [int(a) for a in sys.argv[2:]] or [0]
print("#", rep)
for li in launches:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(li), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    name = rows[0][1] if rows and len(rows[0]) > 1 else "?"
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    seen, body = set(), []
    for r in rows[2:]:
        if len(r) <= ix["# Samples"] or r[0] in seen:
            continue
        seen.add(r[0])
        try:
            body.append((re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]].strip()), int(r[ix["Instructions Executed"]]),
                         int(r[ix["# Samples"]]), {s: int(r[ix[s]] or 0) for s in stalls}))
        except ValueError:
            pass
    tot = sum(b[2] for b in body)
    i2fp = [b[1] for b in body if b[0].startswith("I2FP") and b[1] > 0]
    E = collections.Counter(i2fp).most_common(1)[0][0] if i2fp else 0
    wait = sum(b[2] for b in body if b[0].startswith(("SYNCS", "BAR", "WARPSYNC", "NANOSLEEP")) or
               (b[0].startswith("BRA") and b[1] > 4 * max(E, 1)) or b[0].startswith("EXIT"))
    hot = [b for b in body if E and 0.45 * E <= b[1] <= 2.2 * E and not b[0].startswith(("SYNCS", "EXIT"))]
    hs = sum(b[2] for b in hot)
    mix = collections.Counter()
    for b in hot:
        for k, v in b[3].items():
            mix[k[6:]] += v
    ops = collections.Counter()
    for b in hot:
        ops[b[0].split()[0].split(".")[0]] += round(b[1] / E)
    print(f"\n== launch {li}: {name[:60]}")
    print(f"samples {tot}; barrier / mbarrier waits and exits {100 * wait / max(tot, 1):.0f} %; epilogue pass body {100 * hs / max(tot, 1):.0f} % "
          f"({len(hot)} instructions, executed {E} times each = warp passes)")
    print("stall mix inside the pass body:", ", ".join(f"{k} {100 * v / max(hs, 1):.0f}%" for k, v in mix.most_common(8)))
    print("instructions per pass:", sum(ops.values()), dict(ops.most_common(14)))
