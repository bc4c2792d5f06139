import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, dlq_b200
from dlq_b200 import synth
mode, start = int(sys.argv[1]), int(sys.argv[2])
ctx = dlq_b200.Context(0)
m = dlq_b200.ResNet18(ctx, synth.make_weights(0), synth.load_act_scales(0), 64)
m.set_option("chain_launch_mode", mode); m.set_option("chain_start", start)
x = torch.from_numpy(np.tile(synth.make_input(0, 8), (8, 1, 1, 1))).cuda()
dl = torch.empty((64, 1000), dtype=torch.float32, device="cuda")
for _ in range(2): m.forward(x, dl)
ctx.sync()
print("ok mode", mode, "start", start, "launch mode used", m.plan_info(64, "chain_launch_mode"), "top1", int(dl[0].argmax()))
