run() { echo "== $1"; env $1 timeout 120 python tools/conv_sweep.py --iters 10 --only "$2" 2>&1 >/dev/null | grep "us " | cut -c1-60; }
run X=1 layer1
run DLQ_DBG_PAIR64=1 layer1
run X=1 conv1
run DLQ_DBG_PAIR64=2 conv1
run X=1 .1.conv1
run "DLQ_DBG_B_CAP=12 DLQ_DBG_A_CAP=3" .1.conv1
run "DLQ_DBG_B_CAP=16 DLQ_DBG_A_CAP=2" .1.conv1
run "DLQ_DBG_B_CAP=4 DLQ_DBG_A_CAP=4" .1.conv1
run "DLQ_DBG_B_CAP=6 DLQ_DBG_A_CAP=4" .1.conv1
