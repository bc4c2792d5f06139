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


@pytest.mark.parametrize("n", [1, 3, 17, 40])
def test_network_matches_oracle(ctx, model256, n):
    """n <= 39 runs the fused-shortcut plan (conv1 + 1x1/s2 in one launch, K-split issuers, tile shapes picked for the
    problem size: one tile per item at n = 1 and 3, two or four at 17); n = 40 the plan the benchmark runs at batch 256 -
    the two conv chains, separate 3x3/s2 and 1x1/s2 convs reading parity-plane tensors - on DISTINCT images, all six int8
    checkpoints and the logits.  (The un-fused plan with one launch per conv: test_conv_chain_equals_separate_launches.)"""
    w = synth.make_weights(0)
    x = synth.make_input(7, n)
    assert len({x[i].tobytes() for i in range(n)}) == n
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


def _tile(a, n):
    return np.concatenate([a] * (n // a.shape[0] + 1), 0)[:n]


def test_forward_host_chunked_batch256_matches_oracle(ctx, model256):
    """the e2e headline path: dlq_resnet18_forward_host at N = 256 (64-image chunks, H2D overlapped with compute) on 32
    distinct images (tiled) - every one of the 256 logit rows bit-equal to the oracle's"""
    import torch
    x32 = synth.make_input(11, 32)
    ref = orc.I8Model(synth.make_weights(0), synth.load_act_scales(0)).forward(x32)["logits"]
    xh = torch.from_numpy(_tile(x32, 256)).pin_memory()
    lh = torch.zeros((256, 1000), dtype=torch.float32).pin_memory()
    model256.forward_host(xh, lh)
    assert np.array_equal(lh.numpy().view(np.uint32), _tile(ref, 256).view(np.uint32))
    # a ragged batch (3 chunks of 64 + 8) through the same entry point
    lh.zero_()
    model256.forward_host(xh[:200], lh[:200])
    assert np.array_equal(lh.numpy()[:200].view(np.uint32), _tile(ref, 200).view(np.uint32))
    assert not lh.numpy()[200:].any()


def test_forward_host_u8_chunked_batch256_matches_oracle(ctx, model256):
    """dlq_resnet18_forward_host_u8 at N = 256 (128-image chunks): uint8 HWC images -> device-side table -> network, vs the
    oracle run on the reference's numpy preprocessing of the same images (tools/preprocess_to_bin.py:24-33)"""
    import torch
    import dlq_b200
    rng = np.random.default_rng(5)
    u32 = rng.integers(0, 256, (32, 224, 224, 3), dtype=np.uint8)
    mean = np.array(dlq_b200.ResNet18.IMAGENET_MEAN, dtype=np.float32)
    std = np.array(dlq_b200.ResNet18.IMAGENET_STD, dtype=np.float32)
    x = (u32.astype(np.float32) / np.float32(255.0) - mean) / std
    x = np.ascontiguousarray(np.transpose(x, (0, 3, 1, 2)))
    ref = orc.I8Model(synth.make_weights(0), synth.load_act_scales(0)).forward(x)["logits"]
    model256.set_preprocess()
    uh = torch.from_numpy(_tile(u32, 256)).pin_memory()
    lh = torch.zeros((256, 1000), dtype=torch.float32).pin_memory()
    model256.forward_host_u8(uh, lh)
    assert np.array_equal(lh.numpy().view(np.uint32), _tile(ref, 256).view(np.uint32))


def test_submit_wait_pipeline_matches_oracle(ctx, model256):
    """pipelined host entry points (double-buffered staging, whole-batch forwards): five submits of different batches and
    sizes with two in flight, fp32 and uint8 mixed - every result bit-equal to the oracle"""
    import torch
    import dlq_b200
    w, sc = synth.make_weights(0), synth.load_act_scales(0)
    orc_m = orc.I8Model(w, sc)
    xs = [synth.make_input(20 + i, n) for i, n in enumerate((8, 5, 8))]
    refs = [orc_m.forward(x)["logits"] for x in xs]
    rng = np.random.default_rng(9)
    u = rng.integers(0, 256, (6, 224, 224, 3), dtype=np.uint8)
    mean = np.array(dlq_b200.ResNet18.IMAGENET_MEAN, dtype=np.float32)
    std = np.array(dlq_b200.ResNet18.IMAGENET_STD, dtype=np.float32)
    xu = np.ascontiguousarray(np.transpose((u.astype(np.float32) / np.float32(255.0) - mean) / std, (0, 3, 1, 2)))
    ref_u = orc_m.forward(xu)["logits"]
    model256.set_preprocess()
    # batches of 256 / 160 / 256 images (tiled) so that the forwards take long enough to overlap the copies
    sizes = (256, 160, 256)
    hx = [torch.from_numpy(_tile(x, n)).pin_memory() for x, n in zip(xs, sizes)]
    hu = torch.from_numpy(_tile(u, 192)).pin_memory()
    hl = [torch.zeros((n, 1000), dtype=torch.float32).pin_memory() for n in sizes + (192, 256)]
    model256.submit_host(hx[0], hl[0])
    model256.submit_host(hx[1], hl[1])
    model256.wait()
    model256.submit_host_u8(hu, hl[3])
    model256.wait()
    model256.submit_host(hx[2], hl[2])
    model256.wait()
    model256.submit_host(hx[0], hl[4])
    model256.wait()
    model256.wait()
    model256.wait()      # nothing outstanding: returns at once
    for i, n in enumerate(sizes):
        assert np.array_equal(hl[i].numpy().view(np.uint32), _tile(refs[i], n).view(np.uint32)), f"submit {i}"
    assert np.array_equal(hl[3].numpy().view(np.uint32), _tile(ref_u, 192).view(np.uint32))
    assert np.array_equal(hl[4].numpy().view(np.uint32), _tile(refs[0], 256).view(np.uint32))
    # the synchronous entry point drains outstanding submits before it runs
    model256.submit_host(hx[1], hl[1])
    l2 = torch.zeros((8, 1000), dtype=torch.float32).pin_memory()
    model256.forward_host(torch.from_numpy(xs[2]).pin_memory(), l2)
    assert np.array_equal(l2.numpy().view(np.uint32), refs[2].view(np.uint32))


def test_launch_count_and_span_stamps(ctx, model256):
    """dlq_resnet18_launches_for_batch: 20 kernels when the shortcut convs are fused (N <= 39), 23 otherwise; the span
    stamps of a forward are ordered like its launches"""
    import torch
    assert model256.launches == 23
    assert model256.launches_for_batch(8) == 20 and model256.launches_for_batch(16) == 20
    # from batch 40 on: layer1's four convs share one persistent launch, the fifteen convs of layer2..4 another (conv_chain.cuh)
    assert model256.launches_for_batch(17) == 20 and model256.launches_for_batch(39) == 20
    assert model256.launches_for_batch(40) == 6 and model256.launches_for_batch(256) == 6
    assert model256.plan_info(256, "chain_layers") == 19 and model256.plan_info(256, "chains") == 2
    assert model256.plan_info(8, "chain_layers") == 0
    model256.set_option("conv_chain", 0)
    assert model256.launches_for_batch(40) == 23 and model256.launches_for_batch(256) == 23
    model256.set_option("conv_chain", 1)
    model256.set_option("chain_min_batch", 17)
    assert model256.launches_for_batch(17) == 6
    model256.set_option("chain_min_batch", 40)
    x = torch.from_numpy(synth.make_input(0, 40)).cuda()       # (40: every one of the 23 slots is stamped - chain layers included)
    dl = torch.empty((40, 1000), dtype=torch.float32, device="cuda")
    model256.enable_stamps(4)
    for _ in range(3):
        model256.forward(x, dl)
    st = model256.read_stamps()
    model256.enable_stamps(0)
    assert st.shape == (3, 23, 2)
    entry, exit_ = st[..., 0].astype(np.int64), st[..., 1].astype(np.int64)
    assert (exit_ > entry).all() and (exit_ - entry < 5_000_000).all()
    assert (exit_[:, -1] >= exit_[:, :-1].max(axis=1)).all(), "GAP+FC finishes last"
    assert (entry[:, 1:] >= entry[:, :1]).all(), "the input quantisation enters first"
    assert (entry[1:, 0] >= exit_[:-1, 21]).all(), "a forward starts after the previous one's last conv"
    # stamps off again: results unchanged
    model256.forward(x, dl)
    ctx.sync()


@pytest.mark.parametrize("n", [1, 9, 24, 256])
def test_tile_dependency_flags_equal_grid_dependencies(ctx, n):
    """tile-level dependency flags between the conv launches vs grid-level griddepcontrol.wait (the default): identical
    checkpoints and logits; repeated forwards and CUDA-graph replays never change a bit (a race between a consumer
    tile and its producer would); no dependency wait ever times out"""
    import torch
    import dlq_b200
    m = dlq_b200.ResNet18(ctx, synth.make_weights(0), synth.load_act_scales(0), n)
    m.set_option("tile_flags", 1)
    x = _tile(synth.make_input(13, min(n, 32)), n)
    want = _forward(ctx, m, x)
    if n <= 32:
        ref = orc.I8Model(synth.make_weights(0), synth.load_act_scales(0)).forward(x[:min(n, 32)], checkpoints=True)
        for k in CKPTS:
            assert np.array_equal(want[k], ref[k]), k
    m.set_option("tile_flags", 0)
    got = _forward(ctx, m, x)
    for k in list(CKPTS) + ["logits"]:
        assert np.array_equal(got[k], want[k]), k
    m.set_option("tile_flags", 1)
    dx = torch.from_numpy(x).cuda()
    dl = torch.empty((n, 1000), dtype=torch.float32, device="cuda")
    x2 = _tile(synth.make_input(77, min(n, 32)), n)
    want2 = _forward(ctx, m, x2)["logits"]
    dx2 = torch.from_numpy(x2).cuda()
    dl2 = torch.empty((n, 1000), dtype=torch.float32, device="cuda")
    ctx.sync()
    for rep in range(40 if n < 256 else 15):        # two different batches back to back, no host synchronisation
        m.forward(dx, dl)
        m.forward(dx2, dl2)
        if rep % 5 == 4:
            ctx.sync()
            assert np.array_equal(dl.cpu().numpy().view(np.uint32), want["logits"].view(np.uint32)), rep
            assert np.array_equal(dl2.cpu().numpy().view(np.uint32), want2.view(np.uint32)), rep
    m.graph_capture(dx, dl)
    for rep in range(20):
        m.graph_launch()
    ctx.sync()
    assert np.array_equal(dl.cpu().numpy().view(np.uint32), want["logits"].view(np.uint32))
    # a different batch size on the same model (other plan, same counters)
    if n >= 9:
        got2 = _forward(ctx, m, x[:5])
        assert np.array_equal(got2["logits"].view(np.uint32), want["logits"][:5].view(np.uint32))
        got3 = _forward(ctx, m, x)
        assert np.array_equal(got3["layer3"], want["layer3"])
    assert m.dep_timeouts == 0
    m.close()


@pytest.mark.parametrize("n,start", [(17, 7), (40, 7), (256, 7), (64, 19), (33, 13), (24, 11), (48, 8)])
def test_conv_chain_equals_separate_launches(ctx, n, start):
    """the persistent multi-layer kernel (default from batch 40 on) vs one launch per conv: identical checkpoints and
    logits, stable over repeated forwards and graph replays, and - where the oracle finishes in seconds - equal to it"""
    import torch
    import dlq_b200
    m = dlq_b200.ResNet18(ctx, synth.make_weights(0), synth.load_act_scales(0), n)
    m.set_option("chain_min_batch", 17)     # (default 40: these cases exercise the chains on small batches)
    m.set_option("chain_start", start)      # 7 = layer2.0.conv1 (default), 8 = layer2.0.conv2, 19 = layer4.0.conv1, 13 = layer3.0.conv1, ...
    m.set_option("chain_layer1", 0 if start == 13 else 1)
    x = _tile(synth.make_input(17, min(n, 40)), n)
    want = _forward(ctx, m, x)
    assert m.launches_for_batch(n) < 23, "the chain replaces several launches"
    if n <= 40:
        ref = orc.I8Model(synth.make_weights(0), synth.load_act_scales(0)).forward(x, checkpoints=True)
        for k in CKPTS:
            assert np.array_equal(want[k], ref[k]), k
        assert np.array_equal(want["logits"].view(np.uint32), ref["logits"].view(np.uint32))
    m.set_option("conv_chain", 0)
    got = _forward(ctx, m, x)
    assert m.launches_for_batch(n) == (23 if n >= 40 else 20)      # (below 40 the shortcut convs ride on conv1's launch)
    for k in list(CKPTS) + ["logits"]:
        assert np.array_equal(got[k], want[k]), k
    m.set_option("conv_chain", 1)
    dx = torch.from_numpy(x).cuda()
    dl = torch.empty((n, 1000), dtype=torch.float32, device="cuda")
    # back-to-back forwards WITHOUT host synchronisation, alternating between two different batches (and two batch sizes):
    # a consumer tile that ran ahead of its producer would pick up the other batch's activations
    x2 = _tile(synth.make_input(99, min(n, 40)), n)
    want2 = _forward(ctx, m, x2)
    n3 = max(17, n // 2 + 1)
    want3 = _forward(ctx, m, x2[:n3])["logits"]
    dx2 = torch.from_numpy(x2).cuda()
    dl2 = torch.empty((n, 1000), dtype=torch.float32, device="cuda")
    dl3 = torch.empty((n3, 1000), dtype=torch.float32, device="cuda")
    ctx.sync()
    for rep in range(12 if n < 256 else 6):
        m.forward(dx, dl)
        m.forward(dx2[:n3], dl3)
        m.forward(dx2, dl2)
        if rep % 3 == 2:
            ctx.sync()
            assert np.array_equal(dl.cpu().numpy().view(np.uint32), want["logits"].view(np.uint32)), rep
            assert np.array_equal(dl2.cpu().numpy().view(np.uint32), want2["logits"].view(np.uint32)), rep
            assert np.array_equal(dl3.cpu().numpy().view(np.uint32), want3.view(np.uint32)), rep
    m.graph_capture(dx, dl)
    for rep in range(10):
        m.graph_launch()
    ctx.sync()
    assert np.array_equal(dl.cpu().numpy().view(np.uint32), want["logits"].view(np.uint32))
    assert m.dep_timeouts == 0
    m.close()


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
