#!/usr/bin/env python
"""profiles/rNN_traffic.json from a tools/ncu_summary.py table: DRAM bytes (read + write) per launch of ONE forward (the
last six launches of the capture), the conv family's total, and the per-kernel figures bench.py's roofline_bw quotes.
    python tools/ncu_traffic.py profiles/r02_step_full_summary.txt > profiles/r02_traffic.json"""
import json
import sys

rows = [l.split() for l in open(sys.argv[1]) if l[:1].isdigit()]
fwd = rows[-6:]
per = []
for r in fwd:
    # idx kernel... grid dur tensor cycles rd wr inst regs warps dram  (the kernel name may contain blanks)
    tail = r[-9:]
    name = " ".join(r[1:-10])
    per.append({"kernel": name, "dur_us": float(tail[0]), "tensor_pct": float(tail[1]),
                "dram_bytes": (float(tail[3]) + float(tail[4])) * 1e6})
bw = {}
for p in per:
    for key, pat in (("quantize_s2d", "stem_s2d"), ("maxpool", "maxpool"), ("gap_fc", "gap_fc")):
        if pat in p["kernel"]:
            bw[key] = p["dram_bytes"]
conv = sum(p["dram_bytes"] for p in per if "conv_" in p["kernel"])
print(json.dumps({"dram_bytes_per_launch": bw, "per_launch": per, "conv_family_dram_bytes_per_step": conv,
                  "source": f"{sys.argv[1]} (last six launches = one forward): ncu --set full --clock-control none of `bench.py --steps 2 "
                            "--warmup 3 --no-extras --no-cpu-baseline --chain-launch-mode 2` (tools/gpu_ncu.sh), dram__bytes_read.sum + "
                            "dram__bytes_write.sum per launch; caches flushed between launches (no inter-kernel L2 reuse)"}, indent=1))
