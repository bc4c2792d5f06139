"""CPU: pins the oracle's FP32 half against (a) the reference's own committed known-answer vectors
(RKL/tmp_e2e -> RKL/out/step8_logits.bin, copied to tests/golden/ref_gapfc_*.bin), (b) outputs of the reference's
own MN/v3.c forward (tests/golden/mnist_v3_seed.npz, generated from oracle/_ref/libref_mnist_v3.so),
(c) PyTorch / torchvision — the third-party arithmetic every reference test compares with (atol 1e-4)."""
import os

import numpy as np
import pytest

import orc
from dlq_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _bin(name, shape):
    return np.fromfile(os.path.join(GOLD, name), dtype=np.float32).reshape(shape)


def test_reference_gap_known_answer():
    l4, gap = _bin("ref_gapfc_l4.bin", (1, 512, 7, 7)), _bin("ref_gapfc_gap.bin", (1, 512))
    got = orc.gap_f32(l4)
    assert np.abs(got - gap).max() <= 1e-6        # SURVEY §4 measured 4.8e-7 with a plain mean


def test_reference_fc_known_answer():
    gap = _bin("ref_gapfc_gap.bin", (1, 512))
    w, b = _bin("ref_gapfc_fc_weight.bin", (1000, 512)), _bin("ref_gapfc_fc_bias.bin", (1000,))
    ref = _bin("ref_gapfc_logits.bin", (1, 1000))
    got = orc.fc_f32(gap, w, b)
    assert np.abs(got - ref).max() <= 1e-4        # the reference's own criterion (R/infer_head.cu:125-127)
    assert int(got.argmax()) == int(ref.argmax()) == 293


def test_mnist_forward_matches_reference_v3():
    g = np.load(os.path.join(GOLD, "mnist_v3_seed.npz"))
    hidden, out = orc.mnist_forward(g["x"], g["w1"], g["b1"], g["w2"], g["b2"])
    # same C arithmetic, same summation order; only expf may differ in the last ulp between libm builds
    assert np.abs(hidden - g["hidden"]).max() <= 1e-5
    assert np.abs(out - g["out"]).max() <= 1e-6
    assert np.array_equal(out.argmax(1), g["out"].argmax(1))


def test_fp32_ops_vs_torch():
    torch = pytest.importorskip("torch")
    import torch.nn.functional as F
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 8, 13, 11)).astype(np.float32)
    w = (rng.standard_normal((6, 8, 3, 3)) * 0.2).astype(np.float32)
    for stride, pad in [(1, 1), (2, 1), (2, 0), (1, 0)]:
        ref = F.conv2d(torch.from_numpy(x), torch.from_numpy(w), stride=stride, padding=pad).numpy()
        assert np.abs(orc.conv2d_f32(x, w, stride, pad) - ref).max() <= 1e-4
    w7 = (rng.standard_normal((4, 8, 7, 7)) * 0.05).astype(np.float32)
    ref = F.conv2d(torch.from_numpy(x), torch.from_numpy(w7), stride=2, padding=3).numpy()
    assert np.abs(orc.conv2d_f32(x, w7, 2, 3) - ref).max() <= 1e-4
    g, b, m = (rng.standard_normal(8).astype(np.float32) for _ in range(3))
    v = rng.uniform(0.5, 1.5, 8).astype(np.float32)
    ref = F.batch_norm(torch.from_numpy(x), torch.from_numpy(m), torch.from_numpy(v), torch.from_numpy(g),
                       torch.from_numpy(b), False, 0.0, 1e-5).numpy()
    assert np.abs(orc.bn_inference_f32(x, g, b, m, v) - ref).max() <= 1e-5
    assert np.array_equal(orc.relu_f32(x), np.maximum(x, 0))
    assert np.array_equal(orc.add_f32(x, x[::-1].copy()), x + x[::-1])
    ref = F.max_pool2d(torch.from_numpy(x), 3, 2, 1).numpy()
    assert np.array_equal(orc.maxpool_f32(x), ref)
    assert np.abs(orc.gap_f32(x) - x.mean((2, 3))).max() <= 1e-6
    a, bb = rng.standard_normal((5, 40)).astype(np.float32), rng.standard_normal((40, 7)).astype(np.float32)
    assert np.abs(orc.sgemm_f32(a, bb) - a @ bb).max() <= 1e-4
    lg = rng.standard_normal(1000).astype(np.float32)
    assert np.abs(orc.softmax_f32(lg) - F.softmax(torch.from_numpy(lg), 0).numpy()).max() <= 1e-6


def test_fp32_network_vs_torchvision():
    """whole FP32 oracle network vs torchvision.models.resnet18 carrying the same synthetic weights
    (the reference's step-8 check: R/infer_e2e.cu dumps vs tools/make_e2e_fixtures.py, atol 1e-4 scaled)."""
    torch = pytest.importorskip("torch")
    tv = pytest.importorskip("torchvision")
    w = synth.make_weights(1, fill=orc.fill_f32)
    net = tv.models.resnet18(weights=None).eval()
    sd = net.state_dict()
    for k in sd:
        if k in w:
            sd[k] = torch.from_numpy(w[k])
    net.load_state_dict(sd)
    x = synth.make_input(5, 1, fill=orc.fill_f32)
    feats = {}
    with torch.no_grad():
        t = torch.from_numpy(x)
        t = net.maxpool(net.relu(net.bn1(net.conv1(t))))
        feats["stem_pool"] = t.numpy()
        for i, layer in enumerate((net.layer1, net.layer2, net.layer3, net.layer4), 1):
            t = layer(t)
            feats[f"layer{i}"] = t.numpy()
        t = torch.flatten(net.avgpool(t), 1)
        feats["gap"] = t.numpy()
        feats["logits"] = net.fc(t).numpy()
    got = orc.F32Model(w).forward(x, checkpoints=True)
    for k, ref in feats.items():
        scale = max(1.0, float(np.abs(ref).max()))
        err = float(np.abs(got[k] - ref).max()) / scale
        assert err <= 1e-4, (k, err)
    assert int(got["logits"].argmax()) == int(feats["logits"].argmax())
