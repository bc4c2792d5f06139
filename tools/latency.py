#!/usr/bin/env python
"""BASELINE config 2: ResNet-18 INT8, batch 1, 224x224, one B200 - latency of one forward.
Device-resident fp32 input, one CUDA-graph replay per inference (dlq_resnet18_graph_capture / _launch), each
replay bracketed by CUDA events; reports median / p99 / mean over --iters replays, the plain stream-launch latency,
and checks the graph's logits against the stream path (bit-exact).  Prints one JSON object."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def measure(batch: int, iters: int, chain_min_batch=None):
    import torch
    import dlq_b200
    from dlq_b200 import synth
    ctx = dlq_b200.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream)
    m = dlq_b200.ResNet18(ctx, synth.make_weights(0), synth.load_act_scales(0), batch)
    if chain_min_batch is not None:
        m.set_option("chain_min_batch", chain_min_batch)
    x = torch.from_numpy(synth.make_input(0, batch)).cuda()
    ref = torch.empty((batch, 1000), dtype=torch.float32, device="cuda")
    out = torch.zeros((batch, 1000), dtype=torch.float32, device="cuda")
    m.forward(x, ref)
    ctx.sync()
    m.graph_capture(x, out)
    for _ in range(20):
        m.graph_launch()
    ctx.sync()
    same = bool(torch.equal(out, ref))
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(stream)
        m.graph_launch()
        b.record(stream)
    ctx.sync()
    g = np.array([a.elapsed_time(b) for a, b in ev]) * 1e3
    for a, b in ev:
        a.record(stream)
        m.forward(x, out)
        b.record(stream)
    ctx.sync()
    s = np.array([a.elapsed_time(b) for a, b in ev]) * 1e3
    m.close()
    ctx.close()
    q = lambda v: {"median_us": float(np.median(v)), "p99_us": float(np.percentile(v, 99)), "mean_us": float(v.mean())}
    return {"batch": batch, "iters": iters, "chain_min_batch": chain_min_batch, "graph": q(g), "stream_launches": q(s), "graph_equals_stream": same,
            "images_per_s_graph": batch / (np.median(g) * 1e-6)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, nargs="+", default=[1])
    ap.add_argument("--iters", type=int, default=1000)
    ap.add_argument("--chain-min-batch", type=int, default=None, help="plan batches >= this with the persistent conv chains")
    a = ap.parse_args()
    print(json.dumps({"config": "ResNet-18 INT8 latency, device-resident input, CUDA graph replay",
                      "results": [measure(b, a.iters, a.chain_min_batch) for b in a.batch]}))
