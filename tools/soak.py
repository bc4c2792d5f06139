#!/usr/bin/env python
"""Soak of the persistent conv chains (tile-level dependency flags): thousands of forwards, two batch sizes and two
different inputs alternating back to back without host synchronisation; every checked forward's logits must be the bits
of the first one with the same input, and no dependency wait may time out.
    python tools/soak.py [--forwards 6000]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--forwards", type=int, default=6000)
    a = ap.parse_args()
    import torch
    import dlq_b200
    from dlq_b200 import synth
    ctx = dlq_b200.Context(0)
    w, s = synth.make_weights(0), synth.load_act_scales(0)
    m = dlq_b200.ResNet18(ctx, w, s, 256)
    base = synth.make_input(0, 8)
    cases = []
    for seed, n in ((0, 256), (1, 40), (2, 256), (3, 17)):
        x = torch.from_numpy(np.ascontiguousarray(np.roll(np.tile(base, (32, 1, 1, 1)), seed, axis=0)[:n])).cuda()
        out = torch.empty((n, 1000), dtype=torch.float32, device="cuda")
        m.forward(x, out)
        ctx.sync()
        cases.append((x, out, out.clone()))
    bad = 0
    for i in range(a.forwards):
        x, out, ref = cases[i % len(cases)]
        m.forward(x, out)
        if i % 97 == 0:
            ctx.sync()
            if not torch.equal(out, ref):
                bad += 1
    ctx.sync()
    for x, out, ref in cases:
        bad += 0 if torch.equal(out, ref) else 1
    print({"forwards": a.forwards, "mismatches": bad, "dep_timeouts": m.dep_timeouts})
    m.close()
    ctx.close()
    sys.exit(1 if bad or m is None else 0)


if __name__ == "__main__":
    main()
