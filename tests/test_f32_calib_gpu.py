"""GPU parity of the FP32 reference-semantics network (dlq_resnet18_f32_*), the PTQ calibrator built on it, the
weight-directory loader and the accuracy-harness operators, all through the C ABI, against
the CPU oracle's FP32 half - which is itself pinned to the reference's kernels run live (test_reference_live_gpu.py)
and to the reference's GAP+FC known-answer vectors (test_oracle_golden.py)."""
import numpy as np
import pytest

import orc
from dlq_b200 import synth

pytestmark = pytest.mark.gpu
CKPTS = {"stem_pool": (64, 56, 56), "layer1": (64, 56, 56), "layer2": (128, 28, 28), "layer3": (256, 14, 14),
         "layer4": (512, 7, 7), "gap": (512,)}


def _f32_forward(ctx, m, x):
    import torch
    dx = torch.from_numpy(x).cuda()
    dl = torch.empty((x.shape[0], 1000), dtype=torch.float32, device="cuda")
    m.forward(dx, dl)
    ctx.sync()
    out = {"logits": dl.cpu().numpy()}
    for k, shp in CKPTS.items():
        t = torch.empty((x.shape[0],) + shp, dtype=torch.float32, device="cuda")
        m.checkpoint(k, t)
        ctx.sync()
        out[k] = t.cpu().numpy()
    return out


def test_f32_network_bit_exact_vs_oracle(ctx):
    """every checkpoint, the logits and the recorded absmax equal the oracle's FP32 restatement bit for bit"""
    import dlq_b200
    w = synth.make_weights(0)
    x = synth.make_input(5, 3)
    am_ref = np.zeros(orc.NUM_ACTS, dtype=np.float32)
    ref = orc.F32Model(w).forward(x, checkpoints=True, absmax=am_ref)
    m = dlq_b200.ResNet18F32(ctx, w, 4)
    got = _f32_forward(ctx, m, x)
    for k in list(CKPTS) + ["logits"]:
        bad = int((got[k].view(np.uint32) != ref[k].view(np.uint32)).sum())
        assert bad == 0, f"{k}: {bad}/{ref[k].size} fp32 values differ from the oracle"
    assert np.array_equal(m.absmax().view(np.uint32), am_ref.view(np.uint32))
    m.close()


def test_calibrator_reproduces_committed_scales(ctx):
    """PTQ on the GPU over the calibration batch of tests/golden/make_golden.py == dlq_b200/synth_act_scales.json
    (made by the CPU oracle): same 27 fp32 bit patterns, in two feeds of 4 images and in one feed of 8"""
    import torch
    import dlq_b200
    for seed in (0, 1):
        w = synth.make_weights(seed)
        x = synth.make_input(0, 8)
        m = dlq_b200.ResNet18F32(ctx, w, 8)
        dl = torch.empty((8, 1000), dtype=torch.float32, device="cuda")
        dx = torch.from_numpy(x).cuda()
        m.forward(dx[:4].contiguous(), dl[:4])
        m.forward(dx[4:].contiguous(), dl[4:])
        s_two = m.act_scales()
        m.reset_absmax()
        m.forward(dx, dl)
        s_one = m.act_scales()
        want = np.asarray(synth.load_act_scales(seed), dtype=np.float32)
        assert np.array_equal(s_two.view(np.uint32), want.view(np.uint32))
        assert np.array_equal(s_one.view(np.uint32), want.view(np.uint32))
        # E4M3: the same absmax mapped to 448 (QUANT_SPEC 6), rounded once from double
        am = m.absmax()
        s8 = m.act_scales(fp8=True)
        assert np.array_equal(s8, (np.where(am > 0, am, np.float32(1)).astype(np.float64) / 448.0).astype(np.float32))
        assert np.allclose(s8, orc.fp8_act_scales(want), rtol=2e-7, atol=0)
        m.close()


def test_weight_dir_to_int8_network(ctx, tmp_path):
    """export -> C++ loader -> calibrate on the GPU -> INT8 network: same logits as the in-memory path"""
    import torch
    import dlq_b200
    w = synth.make_weights(0)
    d = str(tmp_path / "w")
    dlq_b200.save_weight_dir(d, w)                 # no scales in the directory
    wd = dlq_b200.WeightDir(d)
    assert not wd.act_scale.any()
    f = dlq_b200.ResNet18F32(ctx, wd, 8)
    dx = torch.from_numpy(synth.make_input(0, 8)).cuda()
    dl = torch.empty((8, 1000), dtype=torch.float32, device="cuda")
    f.forward(dx, dl)
    scales = f.act_scales()
    f.close()
    wd.set_act_scale(scales)
    wd.save(d, with_scales=True)                   # the "quant" block travels with the directory from now on
    wd.close()
    wd2 = dlq_b200.WeightDir(d)
    assert np.array_equal(wd2.act_scale, scales)
    m_dir = dlq_b200.ResNet18(ctx, wd2, None, 4)
    m_mem = dlq_b200.ResNet18(ctx, w, synth.load_act_scales(0), 4)
    x = torch.from_numpy(synth.make_input(3, 4)).cuda()
    l1 = torch.empty((4, 1000), dtype=torch.float32, device="cuda")
    l2 = torch.empty((4, 1000), dtype=torch.float32, device="cuda")
    m_dir.forward(x, l1)
    m_mem.forward(x, l2)
    ctx.sync()
    assert torch.equal(l1, l2)
    m_dir.close(); m_mem.close(); wd2.close()


def test_topk_and_compare(ctx):
    import torch
    rng = np.random.default_rng(0)
    x = rng.standard_normal((37, 1000)).astype(np.float32)
    x[3, 10] = x[3, 500] = 9.0                      # a tie: the lower index wins (reference's strict '>' scan)
    dx = torch.from_numpy(x).cuda()
    idx = torch.empty((37, 5), dtype=torch.int32, device="cuda")
    val = torch.empty((37, 5), dtype=torch.float32, device="cuda")
    ctx.topk_f32(dx, 5, idx, val)
    ctx.sync()
    order = np.lexsort((np.arange(1000)[None, :].repeat(37, 0), -x), axis=1)[:, :5]
    assert np.array_equal(idx.cpu().numpy(), order.astype(np.int32))
    assert np.array_equal(val.cpu().numpy(), np.take_along_axis(x, order, 1))
    assert idx[3, 0].item() == 10 and idx[3, 1].item() == 500
    a = rng.standard_normal(100003).astype(np.float32)
    b = (a + 0.01 * rng.standard_normal(100003)).astype(np.float32)
    r = ctx.compare_f32(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    cos = float(np.dot(a.astype(np.float64), b.astype(np.float64)) / (np.linalg.norm(a.astype(np.float64)) * np.linalg.norm(b.astype(np.float64))))
    assert r["max_abs"] == float(d.max())
    assert abs(r["mean_abs"] - float(d.mean())) < 1e-12
    assert abs(r["cosine"] - cos) < 1e-12


def test_int8_tracks_fp32(ctx):
    """accuracy harness on synthetic data: the INT8 network follows the FP32 one (cosine per checkpoint, top-1)"""
    import torch
    import dlq_b200
    w = synth.make_weights(0)
    n = 8
    x = torch.from_numpy(synth.make_input(11, n)).cuda()
    f = dlq_b200.ResNet18F32(ctx, w, n)
    q = dlq_b200.ResNet18(ctx, w, synth.load_act_scales(0), n)
    lf = torch.empty((n, 1000), dtype=torch.float32, device="cuda")
    lq = torch.empty((n, 1000), dtype=torch.float32, device="cuda")
    f.forward(x, lf)
    q.forward(x, lq)
    r = ctx.compare_f32(lf, lq)
    assert r["cosine"] > 0.99, r
    tf = torch.empty((n, 1), dtype=torch.int32, device="cuda")
    tq = torch.empty((n, 1), dtype=torch.int32, device="cuda")
    ctx.topk_f32(lf, 1, tf)
    ctx.topk_f32(lq, 1, tq)
    ctx.sync()
    assert (tf == tq).float().mean().item() >= 0.75
    f.close(); q.close()


def test_error_codes(ctx):
    """bad arguments return 1 with a message (the reference's usage / IO exit code), never abort"""
    import torch
    import dlq_b200
    w = synth.make_weights(0)
    f = dlq_b200.ResNet18F32(ctx, w, 2)
    x = torch.zeros((3, 3, 224, 224), dtype=torch.float32, device="cuda")
    l = torch.zeros((3, 1000), dtype=torch.float32, device="cuda")
    with pytest.raises(dlq_b200.DlqError) as e:
        f.forward(x, l)                                   # batch 3 > max_batch 2
    assert e.value.code == 1 and "max_batch" in str(e.value)
    with pytest.raises(dlq_b200.DlqError) as e:
        f.checkpoint("layer1", l)                         # no forward has run yet
    assert e.value.code == 1
    f.forward(x[:2].contiguous(), l[:2])
    with pytest.raises(dlq_b200.DlqError) as e:
        f.checkpoint("layer9", l)
    assert e.value.code == 1 and "unknown checkpoint" in str(e.value)
    idx = torch.zeros((3, 40), dtype=torch.int32, device="cuda")
    with pytest.raises(dlq_b200.DlqError) as e:
        ctx.topk_f32(l, 40, idx)                          # k > 32
    assert e.value.code == 1
    bad = dict(w)
    del bad["layer2.0.downsample.0.weight"]
    with pytest.raises((dlq_b200.DlqError, KeyError)):
        dlq_b200.ResNet18F32(ctx, bad, 2)
    f.close()
