#!/bin/bash
timeout 150 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 100 python tools/profile_layers.py 1 2>&1 | tail -1
timeout 100 python tools/latency.py --batch 1 --iters 300 2>&1 | tail -1 | cut -c1-400
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench_last.json 2> gpurun_out/bench_err.log < /dev/null
python tools/benchsum.py gpurun_out/bench_last.json
