#!/bin/bash
# Tuning: what bounds the epilogue-bound layers?  The TIMING build (make TIMING=1 -> libdlq_b200_timing.so) with parts
# of the epilogue switched off (DLQ_DBG_FLAGS: 1 = no epilogue at all, 16 = no alpha/beta shared-memory reads, 32 = no
# staging / global stores); results
# are wrong with any flag set, only the cycle counters matter.
export DLQ_B200_LIB=$PWD/dlq_b200/libdlq_b200_timing.so
for fl in ${FLAGS:-0 1 16 32 48}; do
  for l in ${LAYERS:-conv1 layer1.conv1 layer2.0.downsample}; do
    echo "== flags $fl $l"
    DLQ_DBG_FLAGS=$fl DLQ_DBG_TIMES=1 timeout 60 python tools/conv_sweep.py --iters 2 --only $l --exact 2>&1 >/dev/null | grep "dbg_times" | tail -1 | cut -c1-330
  done
done
