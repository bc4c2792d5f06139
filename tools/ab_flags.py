#!/usr/bin/env python
"""A/B in ONE process on ONE box: tile-level dependency flags (default) vs grid-level dependencies, batch B.
Prints ms/step of both, alternating several times, plus the per-launch in-step spans of each mode."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--rounds", type=int, default=3)
    args = ap.parse_args()
    import torch
    import dlq_b200
    import bench
    from dlq_b200 import synth
    ctx = dlq_b200.Context(0)
    B = args.batch
    m = dlq_b200.ResNet18(ctx, synth.make_weights(0), synth.load_act_scales(0), B)
    x = torch.from_numpy(np.tile(synth.make_input(0, 8), (B // 8 + 1, 1, 1, 1))[:B]).cuda()
    dl = torch.empty((B, 1000), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    ctx.auto_order = False
    stream = torch.cuda.ExternalStream(ctx.stream)

    def run(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(5):
            m.forward(x, dl)
        ctx.sync()
        e0.record(stream)
        for _ in range(steps):
            m.forward(x, dl)
        e1.record(stream)
        ctx.sync()
        return e0.elapsed_time(e1) / steps

    res = {0: [], 1: []}
    for _ in range(args.rounds):
        for mode in (1, 0):
            m.set_option("tile_flags", mode)
            res[mode].append(run(args.steps))
    for mode in (1, 0):
        print(f"tile_flags={mode}: ms/step {['%.4f' % v for v in res[mode]]}  best {min(res[mode]):.4f}")
    names = m.LAUNCH_NAMES
    for mode in (1, 0):
        m.set_option("tile_flags", mode)
        m.enable_stamps(12)
        run(7)
        sp = bench.in_step_spans(m.read_stamps(), names)
        m.enable_stamps(0)
        rep = sorted(sp[2:], key=lambda d: d["conv_union_ms"])[len(sp[2:]) // 2]
        print(f"-- tile_flags={mode}: conv union {rep['conv_union_ms']:.4f} ms, sum {rep['conv_sum_ms']:.4f}")
        print("   span     " + "  ".join(f"{n.replace('layer', 'L')}:{b - a:.1f}" for n, a, b in rep["spans_us"]))
        ends = [b for _, _, b in rep["spans_us"]]
        print("   marginal " + "  ".join(f"{n.replace('layer', 'L')}:{b - (ends[i - 1] if i else 0):.1f}"
                                          for i, (n, a, b) in enumerate(rep["spans_us"])))
    print("dep_timeouts", m.dep_timeouts)


if __name__ == "__main__":
    main()
