#!/usr/bin/env python
"""Per-launch device time of one forward at batch B (dlq_resnet18_profile): python tools/profile_layers.py 1 256"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dlq_b200
from dlq_b200 import synth
ctx = dlq_b200.Context(0)
for B in [int(a) for a in sys.argv[1:]] or [1]:
    m = dlq_b200.ResNet18(ctx, synth.make_weights(0), synth.load_act_scales(0), B)
    x = torch.from_numpy(synth.make_input(0, min(B, 4))).cuda()
    if B > 4:
        x = x.repeat((B + 3) // 4, 1, 1, 1)[:B].contiguous()
    out = torch.empty((B, 1000), dtype=torch.float32, device="cuda")
    for _ in range(3):
        m.forward(x, out)
    ctx.sync()
    acc = np.zeros(m.launches)
    for _ in range(10):
        acc += m.profile(x, out)
    acc /= 10
    print(f"B={B} total {acc.sum()*1e3:.0f} us: " + " ".join(f"{n.replace('layer','L').replace('downsample','ds').replace('conv','c')}={v*1e3:.1f}" for n, v in zip(m.LAUNCH_NAMES, acc)))
    m.close()
