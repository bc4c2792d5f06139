#!/usr/bin/env python
"""One INT8 and one E4M3 forward at a chain-planned batch (17) and at a small one (3), for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize.py
prints the top-1 classes; any out-of-bounds or misaligned access of the kernels shows up in the sanitizer's report."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import dlq_b200
    from dlq_b200 import synth
    ctx = dlq_b200.Context(0)
    w, s = synth.make_weights(0), synth.load_act_scales(0)
    for fp8 in (False, True):
        sc = (np.asarray(s, dtype=np.float64) * 127.0 / 448.0).astype(np.float32) if fp8 else s
        m = dlq_b200.ResNet18(ctx, w, sc, 17, fp8=fp8)
        for n in (17, 3):
            x = torch.from_numpy(synth.make_input(0, n)).cuda()
            out = torch.empty((n, 1000), dtype=torch.float32, device="cuda")
            m.forward(x, out)
            ctx.sync()
            print("fp8" if fp8 else "int8", n, out.argmax(1)[:4].tolist(), "dep_timeouts", m.dep_timeouts, flush=True)
        m.close()
    ctx.close()


if __name__ == "__main__":
    main()
