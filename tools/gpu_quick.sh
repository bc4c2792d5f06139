timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 200 python bench.py --no-cpu-baseline --no-extras > gpurun_out/bench_last.json 2> gpurun_out/bench_err.log < /dev/null; python tools/benchsum.py gpurun_out/bench_last.json 2>/dev/null | head -2
