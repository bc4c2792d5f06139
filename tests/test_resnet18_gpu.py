"""GPU parity of the whole INT8 ResNet-18 path (dlq_resnet18_forward through the C ABI) against the CPU oracle
and the committed golden vector; size-independent properties at BASELINE's batch 256."""
import hashlib
import json
import os

import numpy as np
import pytest

import orc
from dlq_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CKPTS = {"stem_pool": (64, 56, 56), "layer1": (64, 56, 56), "layer2": (128, 28, 28), "layer3": (256, 14, 14),
         "layer4": (512, 7, 7), "gap": (512,)}


@pytest.fixture(scope="module")
def model256(ctx):
    import dlq_b200
    w = synth.make_weights(0)
    m = dlq_b200.ResNet18(ctx, w, synth.load_act_scales(0), 256)
    yield m
    m.close()


def _forward(ctx, m, x):
    import torch
    dx = torch.from_numpy(x).cuda()
    dl = torch.empty((x.shape[0], 1000), dtype=torch.float32, device="cuda")
    m.forward(dx, dl)
    ctx.sync()
    out = {"logits": dl.cpu().numpy()}
    for k, shp in CKPTS.items():
        t = torch.empty((x.shape[0],) + shp, dtype=torch.int8, device="cuda")
        m.checkpoint(k, t)
        ctx.sync()
        out[k] = t.cpu().numpy()
    return out


@pytest.mark.parametrize("n", [1, 3])
def test_network_matches_oracle(ctx, model256, n):
    w = synth.make_weights(0)
    x = synth.make_input(7, n)
    ref = orc.I8Model(w, synth.load_act_scales(0)).forward(x, checkpoints=True)
    got = _forward(ctx, model256, x)
    for k in CKPTS:
        bad = int((got[k] != ref[k]).sum())
        assert bad == 0, f"checkpoint {k}: {bad}/{ref[k].size} int8 values differ"
    # logits are bit-exact vs the INT8 oracle by construction (QUANT_SPEC §5)
    assert np.array_equal(got["logits"].view(np.uint32), ref["logits"].view(np.uint32))


def test_network_matches_committed_golden(ctx, model256):
    g = np.load(os.path.join(GOLD, "i8_seed0_n2.npz"))
    x = synth.make_input(0, 2)
    got = _forward(ctx, model256, x)
    assert np.array_equal(got["logits"].view(np.uint32), g["logits"].view(np.uint32))
    dig = json.loads(str(g["digests"]))
    for k in CKPTS:
        assert hashlib.sha256(got[k].tobytes()).hexdigest() == dig[k], k


def test_batch256_is_image_independent(ctx, model256):
    """Full BASELINE batch: every image's logits must equal what the same image produces alone
    (no cross-image leakage through tile boundaries, pad rows or the tile scheduler)."""
    g = np.load(os.path.join(GOLD, "i8_seed0_n2.npz"))
    x2 = synth.make_input(0, 2)
    x = np.concatenate([x2[(i * 7 + i // 3) % 2][None] for i in range(256)], 0)
    got = _forward(ctx, model256, x)
    for i in range(256):
        j = (i * 7 + i // 3) % 2
        assert np.array_equal(got["logits"][i].view(np.uint32), g["logits"][j].view(np.uint32)), f"image {i}"
    # a second forward with a different batch size must not be polluted by the first (pad rows stay zero)
    got2 = _forward(ctx, model256, x2)
    assert np.array_equal(got2["logits"].view(np.uint32), g["logits"].view(np.uint32))


def test_batch1024_is_image_independent(ctx):
    """BASELINE config 4 per-GPU batch (2048 images over 2 GPUs): parity-plane tensors taller than 32768 rows, image
    positions past 2^23 - every image's logits still equal what the image produces alone"""
    import torch
    import dlq_b200
    g = np.load(os.path.join(GOLD, "i8_seed0_n2.npz"))
    x2 = torch.from_numpy(synth.make_input(0, 2)).cuda()
    n = 1024
    pick = torch.tensor([(i * 5 + i // 7) % 2 for i in range(n)], device="cuda")
    x = x2[pick].contiguous()
    m = dlq_b200.ResNet18(ctx, synth.make_weights(0), synth.load_act_scales(0), n)
    dl = torch.empty((n, 1000), dtype=torch.float32, device="cuda")
    m.forward(x, dl)
    ctx.sync()
    want = torch.from_numpy(g["logits"]).cuda()[pick]
    assert torch.equal(dl.view(torch.int32), want.view(torch.int32))
    m.close()
    del x, dl
    torch.cuda.empty_cache()


def test_forward_host_and_quantised_accuracy(ctx, model256):
    """Host-buffer entry point; and the INT8 logits track the FP32 oracle (reported tolerance: cosine)."""
    import torch
    x = synth.make_input(3, 2)
    xh = torch.from_numpy(x).pin_memory()
    lh = torch.empty((2, 1000), dtype=torch.float32).pin_memory()
    model256.forward_host(xh, lh)
    w = synth.make_weights(0)
    ref = orc.I8Model(w, synth.load_act_scales(0)).forward(x)
    assert np.array_equal(lh.numpy().view(np.uint32), ref["logits"].view(np.uint32))
    f = orc.F32Model(w).forward(x)["logits"]
    for i in range(2):
        cos = float(np.dot(f[i], lh.numpy()[i]) / (np.linalg.norm(f[i]) * np.linalg.norm(lh.numpy()[i])))
        assert cos > 0.98, cos


@pytest.mark.gpu
def test_u8_input_path_equals_reference_preprocessing():
    """uint8 HWC images through the device-side table == the reference's numpy preprocessing
    (tools/preprocess_to_bin.py:24-33) followed by the fp32 entry point, bit for bit; host-buffer variant too."""
    import torch
    import dlq_b200
    from dlq_b200 import synth
    rng = np.random.default_rng(7)
    n = 5
    u8 = rng.integers(0, 256, (n, 224, 224, 3), dtype=np.uint8)
    mean = np.array(dlq_b200.ResNet18.IMAGENET_MEAN, dtype=np.float32)
    std = np.array(dlq_b200.ResNet18.IMAGENET_STD, dtype=np.float32)
    x = u8.astype(np.float32) / 255.0
    x = (x - mean) / std
    x = np.ascontiguousarray(np.transpose(x, (0, 3, 1, 2)))
    ctx = dlq_b200.Context(0)
    m = dlq_b200.ResNet18(ctx, synth.make_weights(0), synth.load_act_scales(0), n)
    m.set_preprocess()
    ref = torch.empty((n, 1000), dtype=torch.float32, device="cuda")
    got = torch.empty((n, 1000), dtype=torch.float32, device="cuda")
    m.forward(torch.from_numpy(x).cuda(), ref)
    m.forward_u8(torch.from_numpy(u8).cuda(), got)
    ctx.sync()
    assert torch.equal(ref, got)
    host = np.empty((n, 1000), dtype=np.float32)
    m.forward_host_u8(u8, host)
    assert np.array_equal(host.view(np.uint32), ref.cpu().numpy().view(np.uint32))
    m.close()
    ctx.close()
