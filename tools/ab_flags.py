#!/usr/bin/env python
"""A/B in ONE process on ONE box (boxes differ by several percent): the conv chain (persistent multi-layer kernel, default),
tile-level dependency flags between separate launches, and plain grid-level dependencies, at batch B.
Prints ms/step of every mode, alternating several rounds, plus the per-launch in-step spans of each mode."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

MODES = {
    "chain": {"conv_chain": 1, "tile_flags": 0, "chain_start": 7, "chain_layer1": 1},
    "chain_no_l1": {"conv_chain": 1, "tile_flags": 0, "chain_start": 7, "chain_layer1": 0},
    "chain13": {"conv_chain": 1, "tile_flags": 0, "chain_start": 8, "chain_layer1": 0},
    "chain_from_L2.1": {"conv_chain": 1, "tile_flags": 0, "chain_first_block": 3},
    "chain_from_L3": {"conv_chain": 1, "tile_flags": 0, "chain_first_block": 4},
    "grid": {"conv_chain": 0, "tile_flags": 0},
    "flags": {"conv_chain": 0, "tile_flags": 1},
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--modes", default="chain,grid")
    ap.add_argument("--fp8", action="store_true")
    args = ap.parse_args()
    import torch
    import dlq_b200
    import bench
    from dlq_b200 import synth
    ctx = dlq_b200.Context(0)
    B = args.batch
    sc = np.asarray(synth.load_act_scales(0), dtype=np.float64)
    sc = (sc * 127.0 / 448.0).astype(np.float32) if args.fp8 else sc.astype(np.float32)
    m = dlq_b200.ResNet18(ctx, synth.make_weights(0), sc, B, fp8=args.fp8)
    x = torch.from_numpy(np.tile(synth.make_input(0, 8), (B // 8 + 1, 1, 1, 1))[:B]).cuda()
    dl = torch.empty((B, 1000), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    ctx.auto_order = False
    stream = torch.cuda.ExternalStream(ctx.stream)

    def set_mode(name):
        opts = dict(MODES[name])
        for k, v in opts.items():
            m.set_option(k, v)

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

    modes = args.modes.split(",")
    res = {k: [] for k in modes}
    logits = {}
    for _ in range(args.rounds):
        for k in modes:
            set_mode(k)
            res[k].append(run(args.steps))
            logits[k] = dl.clone()
    for k in modes:
        print(f"{k:16s}: ms/step {['%.4f' % v for v in res[k]]}  best {min(res[k]):.4f}  launches {m.launches_for_batch(B) if k == modes[-1] else ''}"
              f"  logits==first-mode {bool(torch.equal(logits[k], logits[modes[0]]))}")
    names = m.LAUNCH_NAMES
    for k in modes:
        set_mode(k)
        m.enable_stamps(12)
        run(7)
        sp = bench.in_step_spans(m.read_stamps(), names)
        m.enable_stamps(0)
        rep = sorted(sp[2:], key=lambda d: d["conv_union_ms"])[len(sp[2:]) // 2]
        print(f"-- {k}: conv union {rep['conv_union_ms']:.4f} ms, sum {rep['conv_sum_ms']:.4f}, period {rep.get('period_ms', 0):.4f}")
        print("   span     " + "  ".join(f"{n.replace('layer', 'L')}:{b - a:.1f}" for n, a, b in rep["spans_us"]))
        ends = [b for _, _, b in rep["spans_us"]]
        print("   end      " + "  ".join(f"{n.replace('layer', 'L')}:{b:.1f}" for n, a, b in rep["spans_us"]))
    set_mode(modes[0])
    print("dep_timeouts", m.dep_timeouts, {k: m.plan_info(B, k) for k in ("chain_layers", "chains", "chain_launch_mode", "chain_cta_pairs", "chain_a_stages", "chain_b_stages", "flag_units")})


if __name__ == "__main__":
    main()
