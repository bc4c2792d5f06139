#!/usr/bin/env python
"""In-step spans (globaltimer stamps, bench.py's method) of every kernel of the INT8 and the E4M3 network at one batch:
where the two differ.     python tools/net_spans.py [--batch 256]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    a = ap.parse_args()
    import torch
    import dlq_b200
    from dlq_b200 import synth
    ctx = dlq_b200.Context(0)
    B = a.batch
    weights, scales = synth.make_weights(0), synth.load_act_scales(0)
    x = torch.from_numpy(np.tile(synth.make_input(0, 2), ((B + 1) // 2, 1, 1, 1))[:B].copy()).cuda()
    logits = torch.empty((B, 1000), dtype=torch.float32, device="cuda")
    out = {}
    for name, fp8 in (("int8", False), ("e4m3", True)):
        sc = (np.asarray(scales, dtype=np.float64) * 127.0 / 448.0).astype(np.float32) if fp8 else scales
        m = dlq_b200.ResNet18(ctx, weights, sc, B, fp8=fp8)
        for _ in range(5):
            m.forward(x, logits)
        m.enable_stamps(8)
        for _ in range(8):
            m.forward(x, logits)
        ctx.sync()
        sp = bench.in_step_spans(m.read_stamps(), m.LAUNCH_NAMES)[2:]
        m.enable_stamps(0)
        med = sorted(sp, key=lambda d: d.get("period_ms", 0))[len(sp) // 2]
        out[name] = {n: (s, e) for n, s, e in med["spans_us"]}
        print(name, "period %.1f us" % (med.get("period_ms", 0) * 1e3), "conv union %.1f us" % (med["conv_union_ms"] * 1e3))
        m.close()
    print("%-22s %18s %18s" % ("kernel", "int8 start-end", "e4m3 start-end"))
    for n in out["int8"]:
        i, f = out["int8"][n], out["e4m3"].get(n, (0, 0))
        print("%-22s %8.1f-%8.1f  %8.1f-%8.1f   (%+.1f us longer)" % (n, i[0], i[1], f[0], f[1], (f[1] - f[0]) - (i[1] - i[0])))


if __name__ == "__main__":
    main()
