#!/bin/bash
# BASELINE config 4 through the C++ driver (examples/dlq_bench over dlq_multi_*): 2048 images batch-sharded over G GPUs,
# INT8 and FP8; plus the multi-GPU parity tests and (G == 2) the bench line under torch.distributed.run.
set -u
G=${1:-2}
O=gpurun_out
timeout 300 python -m pytest tests/test_multi_gpu.py tests/test_driver_gpu.py -m gpu -q 2>&1 | tail -3
timeout 300 examples/_build/dlq_bench --gpus $G --batch 2048 --iters 20 --warmup 3 > $O/r02_cpp_bench_${G}gpu_int8.json 2> $O/cpp_${G}_err.log; cat $O/r02_cpp_bench_${G}gpu_int8.json
timeout 300 examples/_build/dlq_bench --gpus $G --batch 2048 --iters 20 --warmup 3 --fp8 > $O/r02_cpp_bench_${G}gpu_fp8.json 2>> $O/cpp_${G}_err.log; cat $O/r02_cpp_bench_${G}gpu_fp8.json
if [ "$G" = "2" ]; then
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > $O/r02_bench_2gpu.json 2> $O/bench2_err.log; tail -c 600 $O/r02_bench_2gpu.json
fi
