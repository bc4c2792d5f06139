#!/usr/bin/env python
"""Per-layer quantised conv sweep (BASELINE.json config #5): every unique ResNet-18 conv shape at batch B,
timed alone through the native-layout C-ABI entry (dlq_conv2d_i8_act, fused epilogue), reported against
min(tensor, HBM) roofline.  Writes one JSON object per layer to stdout and a table to stderr.

    python tools/conv_sweep.py [--batch 256] [--iters 20] [--only NAME]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# name, IC, H, OC, k, stride, pad, residual   (unique shapes of SURVEY Appendix C)
SHAPES = [
    ("conv1", 3, 224, 64, 7, 2, 3, False),
    ("layer1.conv1", 64, 56, 64, 3, 1, 1, False),
    ("layer1.conv2+res", 64, 56, 64, 3, 1, 1, True),
    ("layer2.0.conv1", 64, 56, 128, 3, 2, 1, False),
    ("layer2.0.downsample", 64, 56, 128, 1, 2, 0, False),
    ("layer2.conv2+res", 128, 28, 128, 3, 1, 1, True),
    ("layer2.1.conv1", 128, 28, 128, 3, 1, 1, False),
    ("layer3.0.conv1", 128, 28, 256, 3, 2, 1, False),
    ("layer3.0.downsample", 128, 28, 256, 1, 2, 0, False),
    ("layer3.conv2+res", 256, 14, 256, 3, 1, 1, True),
    ("layer3.1.conv1", 256, 14, 256, 3, 1, 1, False),
    ("layer4.0.conv1", 256, 14, 512, 3, 2, 1, False),
    ("layer4.0.downsample", 256, 14, 512, 1, 2, 0, False),
    ("layer4.conv2+res", 512, 7, 512, 3, 1, 1, True),
    ("layer4.1.conv1", 512, 7, 512, 3, 1, 1, False),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default=None)
    ap.add_argument("--exact", action="store_true")
    args = ap.parse_args()
    import torch
    import dlq_b200
    ctx = dlq_b200.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    tensor_peak = 2 * peaks["bf16_tflops"]      # burst figure: each kernel is timed alone
    hbm_peak = peaks["hbm_gbs"]
    B = args.batch
    rng = np.random.default_rng(0)
    flush = torch.empty(256 << 20, dtype=torch.int8, device="cuda")
    rows = []
    for name, IC, H, OC, k, s, p, has_res in SHAPES:
        if args.only and (args.only != name if args.exact else args.only not in name):
            continue
        OH = (H + 2 * p - k) // s + 1
        wq = rng.integers(-127, 128, (OC, IC, k, k), dtype=np.int8)
        w = ctx.pack_conv_weights_i8(wq, s, p)
        pr = ctx.required_pad_rows(w)
        if IC == 3:
            xbuf, xa = ctx.new_act(B, H // 2, H // 2 + 3, 32, pr)
        else:
            xbuf, xa = ctx.new_act(B, H, H, IC, pr)
        # every conv but the stem reads post-ReLU activations in the network: same range (and hint) here
        if IC == 3:
            xbuf.random_(-128, 127)     # pad rows too: irrelevant for timing
        else:
            xbuf.random_(0, 127)
        ybuf, ya = ctx.new_act(B, OH, OH, OC, 1)
        rbuf, ra = ctx.new_act(B, OH, OH, OC, 1)
        rbuf.random_(-128, 127)
        alpha = torch.full((OC,), 2.0 ** -9, dtype=torch.float32, device="cuda")
        beta = torch.zeros((OC,), dtype=torch.float32, device="cuda")

        run = ctx.conv_plan(xa, w, ya, alpha, beta, ra if has_res else None, 0.5, True)

        for _ in range(3):
            run()
        ctx.sync()
        tot = 0.0
        for _ in range(args.iters):
            # L2 flush between timed iterations (256 MB > 126 MB L2), enqueued on the SAME stream right in front of
            # the first event and without a host synchronisation in between: the conv launch is already queued when
            # the flush retires, so the interval holds the kernel and not ~9 us of idle-GPU launch latency
            # (profiles/README.md: globaltimer spans of the CTAs vs event times)
            with torch.cuda.stream(stream):
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            run()
            e1.record(stream)
            ctx.sync()
            tot += e0.elapsed_time(e1)
        ms = tot / args.iters
        macs = B * OH * OH * OC * IC * k * k
        in_bytes = B * H * H * (4 if IC == 3 else IC)
        byts = in_bytes + B * OH * OH * OC * (2 if has_res else 1) + OC * IC * k * k + 8 * OC
        tops = 2 * macs / (ms * 1e-3) / 1e12
        gbs = byts / (ms * 1e-3) / 1e9
        t_tensor = 2 * macs / (tensor_peak * 1e12) * 1e3
        t_hbm = byts / (hbm_peak * 1e9) * 1e3
        bound = "tensor" if t_tensor >= t_hbm else "hbm"
        frac = max(t_tensor, t_hbm) / ms
        row = {"layer": name, "batch": B, "ms": ms, "TOPS": tops, "pct_tensor_peak": 100 * tops / tensor_peak,
               "GBs": gbs, "pct_hbm_peak": 100 * gbs / hbm_peak, "bound": bound, "roofline_frac": frac,
               "ideal_ms": max(t_tensor, t_hbm)}
        rows.append(row)
        print(json.dumps(row), flush=True)
        print(f"{name:22s} {ms*1e3:8.1f} us  {tops:7.1f} TOPS ({row['pct_tensor_peak']:5.1f}%)  {gbs:7.0f} GB/s "
              f"({row['pct_hbm_peak']:5.1f}%)  bound={bound:6s} frac={frac:.2f}", file=sys.stderr, flush=True)
        run.destroy()
        w.free()
        del xbuf, ybuf, rbuf
    ctx.close()


if __name__ == "__main__":
    main()
