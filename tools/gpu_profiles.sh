#!/bin/bash
# Round evidence (one gpurun call): the full bench line, per-layer sweeps at batch 256 and 1, latency, then the ncu launch
# list and one full capture of every kernel of a step (tools/gpu_ncu.sh).  Every profiled command is first run plainly.
set -u
O=gpurun_out
R=${1:-r02}
timeout 500 python bench.py --steps 50 --warmup 10 > $O/${R}_bench_b256.json 2> $O/bench_err.log < /dev/null || echo "bench failed"
timeout 120 python tools/conv_sweep.py --iters 10 > $O/${R}_conv_sweep.jsonl 2> $O/${R}_conv_sweep.txt
timeout 120 python tools/conv_sweep.py --iters 20 --batch 1 > $O/${R}_conv_sweep_b1.jsonl 2> $O/${R}_conv_sweep_b1.txt
timeout 100 python tools/latency.py --batch 1 8 32 --iters 500 2>/dev/null | tail -1 > $O/${R}_latency.json
bash tools/gpu_ncu.sh $R
