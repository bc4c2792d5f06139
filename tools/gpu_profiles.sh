#!/bin/bash
# Round evidence (one gpurun call): bench line, ncu launch list of the same command, one full ncu capture of every kernel of one
# step (conv launches + chain + bandwidth kernels), per-layer sweeps at batch 256 and 1.  Every profiled command is first run
# plainly (it must exit 0 without ncu).
set -u
O=gpurun_out
R=${1:-r02}
timeout 400 python bench.py --steps 50 --warmup 10 > $O/${R}_bench_b256.json 2> $O/bench_err.log < /dev/null || echo "bench failed"
timeout 100 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/plain.log 2>&1 < /dev/null && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_ncu_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_launch.log 2>&1 < /dev/null
# two whole forwards (6 launches each) behind the warm-up forwards
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"conv_|maxpool_rows|stem_s2d|gap_fc" -s 30 -c 12 -o $O/${R}_step_full -f \
  python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_full.log 2>&1 < /dev/null
timeout 120 python tools/conv_sweep.py --iters 10 > $O/${R}_conv_sweep.jsonl 2> $O/${R}_conv_sweep.txt
timeout 120 python tools/conv_sweep.py --iters 20 --batch 1 > $O/${R}_conv_sweep_b1.jsonl 2> $O/${R}_conv_sweep_b1.txt
timeout 100 python tools/latency.py --batch 1 8 32 --iters 500 2>/dev/null | tail -1 > $O/${R}_latency.json
tail -3 $O/ncu_full.log
