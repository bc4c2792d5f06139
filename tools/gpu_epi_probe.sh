#!/bin/bash
# Tuning: what bounds the 64-channel layers?  The TIMING build (make TIMING=1 -> libdlq_b200_timing.so) with the
# epilogue's shared-memory traffic switched off piece by piece (DLQ_DBG_FLAGS: 16 = no alpha/beta reads, 32 = no
# staging / stores, 1 = no epilogue at all); results are wrong with any flag set, only the times matter.
export DLQ_B200_LIB=$PWD/dlq_b200/libdlq_b200_timing.so
for fl in 0 16 32 48 1; do
  for l in conv1 layer1.conv1 layer1.conv2+res; do
    echo "== flags $fl $l"
    DLQ_DBG_FLAGS=$fl DLQ_DBG_TIMES=1 timeout 60 python tools/conv_sweep.py --iters 3 --only $l --exact 2>&1 >/dev/null | grep -v "^$" | grep "dbg_times\|^ *$l\|us" | tail -3 | cut -c1-330
  done
done
