#!/bin/bash
# parity tests, then one bench.py line (bounded by timeouts; never reads stdin)
timeout 150 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench_last.json 2> gpurun_out/bench_err.log < /dev/null
python tools/benchsum.py gpurun_out/bench_last.json
