#!/bin/bash
# Standard GPU round-trip: parity tests, then the per-layer sweep (plain, and with per-role cycle counters).
timeout 150 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 120 python tools/conv_sweep.py --iters 10 2>&1 >/dev/null | grep -v "^$" > gpurun_out/sweep_plain.txt
DLQ_DBG_TIMES=1 timeout 120 python tools/conv_sweep.py --iters 2 2>&1 >/dev/null | grep -v "^$" > gpurun_out/sweep_dbg0.txt
cat gpurun_out/sweep_plain.txt
