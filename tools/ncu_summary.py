#!/usr/bin/env python
"""Summarise an ncu report (read here, no GPU): per launch duration, tensor-pipe %, DRAM bytes, registers.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.txt"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "dur_us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
        ("sm__cycles_elapsed.max", "cycles"), ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
        ("smsp__inst_executed.sum", "warp_inst"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct")]
print("# " + rep)
print("idx kernel grid " + " ".join(n for _, n in want))
for n, r in enumerate(rows[2:]):
    vals = []
    for k, _ in want:
        vals.append(r[idx[k]] if k in idx else "na")
    print(n, r[idx["Kernel Name"]].split("(")[0].replace("void ", ""), r[idx["Grid Size"]].replace(" ", ""), " ".join(vals))
