for i in 1 2; do
echo "== prefetch (product build)"; timeout 120 python tools/conv_sweep.py --iters 20 --only "1" 2>&1 >/dev/null | grep "us " | cut -c1-60 | head -4
echo "== no prefetch"; cp dlq_b200/libdlq_b200.so /tmp/keep.so; cp dlq_b200/libdlq_np.so dlq_b200/libdlq_b200.so; timeout 120 python tools/conv_sweep.py --iters 20 --only "1" 2>&1 >/dev/null | grep "us " | cut -c1-60 | head -4; cp /tmp/keep.so dlq_b200/libdlq_b200.so
done
