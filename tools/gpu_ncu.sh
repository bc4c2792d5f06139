#!/bin/bash
# ncu evidence of one bench step (one gpurun call; the plain run first, it must exit 0): launch list + full capture.
# --chain-launch-mode 2: ncu (2025.2, driver 580) dies with LaunchFailed on a cooperative cluster launch made through
# cudaLaunchKernelEx (tools/ncu_chain_probe.py: mode 2 profiles, modes 0 / 1 do not), so the captures launch the SAME
# chain kernel without the cooperative attribute - under the profiler every kernel runs alone on the device anyway.
set -u
O=gpurun_out
R=${1:-r02}
timeout 100 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --chain-launch-mode 2 > $O/plain.log 2>&1 < /dev/null && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_ncu_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --chain-launch-mode 2 > $O/ncu_launch.log 2>&1 < /dev/null
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"conv_|maxpool_rows|stem_s2d|gap_fc" -s 30 -c 12 -o $O/${R}_step_full -f \
  python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --chain-launch-mode 2 > $O/ncu_full.log 2>&1 < /dev/null
tail -4 $O/ncu_full.log; tail -2 $O/ncu_launch.log
