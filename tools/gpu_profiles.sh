#!/bin/bash
# Round evidence: bench line, ncu launch list of the same command, full ncu capture of the conv launches of one step,
# per-layer sweep, probes.  Every profiled command is first run plainly (it must exit 0 without ncu).
set -u
O=gpurun_out
timeout 250 python bench.py --steps 50 --warmup 10 > $O/r01_bench_b256.json 2> $O/bench_err.log < /dev/null || echo "bench failed"
timeout 100 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1 < /dev/null && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r01_ncu_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_launch.log 2>&1 < /dev/null
timeout 400 ncu --set full --import-source on --clock-control none -k regex:conv_i8 -s 60 -c 20 -o $O/r01_conv_full -f \
  python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_full.log 2>&1 < /dev/null
timeout 120 python tools/conv_sweep.py --iters 10 > $O/r01_conv_sweep.jsonl 2> $O/r01_conv_sweep.txt
timeout 60 probe/_build/mma_rate > $O/r01_mma_rate.log 2>&1
timeout 60 probe/_build/ldc_rate > $O/r01_ldc_rate.log 2>&1
timeout 100 python tools/latency.py --batch 1 8 32 --iters 500 2>/dev/null | tail -1 > $O/r01_latency.json
python tools/benchsum.py $O/r01_bench_b256.json
tail -3 $O/ncu_full.log
