"""MNIST MLP forward on the B200 (dlq_b200/mnist.py) against (a) the same INT8 pipeline restated with the
oracle's operators — bit-exact logits — and (b) the reference's own CPU forward (MN/v3.c forward_timed, golden
tests/golden/mnist_v3_seed.npz) — FP32 vs INT8, so by tolerance: arg-max equal, logits within 5 % of their range."""
import os

import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_mnist_mlp_int8_vs_oracle_and_reference():
    import torch
    import dlq_b200
    from dlq_b200.mnist import MnistMLP
    g = np.load(os.path.join(GOLD, "mnist_v3_seed.npz"))
    x, w1, b1, w2, b2 = (g[k] for k in ("x", "w1", "b1", "w2", "b2"))
    ctx = dlq_b200.Context(0)
    m = MnistMLP(ctx, w1, b1, w2, b2, x)
    z, p = m.forward(torch.from_numpy(x).cuda())
    ctx.sync()
    z, p = z.cpu().numpy(), p.cpu().numpy()
    # (a) oracle restatement of the same arithmetic
    xq = orc.quantize(x, float(m.s_x))
    _, h = orc.fc_i8(xq, m.w1q, m.sc1, m.b1)
    hq = orc.quantize(orc.relu_f32(h), float(m.s_h))
    _, z_ref = orc.fc_i8(hq, m.w2q, m.sc2, m.b2)
    assert np.array_equal(z.view(np.uint32), z_ref.view(np.uint32))
    assert np.allclose(p, np.stack([orc.softmax_f32(r) for r in z_ref]), atol=2e-6)
    # (b) the reference's FP32 CPU forward (pre-softmax logits are not stored by v3.c; compare its probabilities)
    ref_p = g["out"]
    assert np.array_equal(p.argmax(1), ref_p.argmax(1))
    assert np.abs(p - ref_p).max() < 0.05
    ctx.close()
