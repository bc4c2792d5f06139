"""GPU parity of the bandwidth-bound operators (quantise / dequantise / BN / ReLU / add / max-pool / GAP / FC /
softmax) through the C ABI vs the CPU oracle; edge cases: empty, ragged (non-multiple-of-16) sizes."""
import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("n", [0, 1, 15, 16, 4097, 3 * 224 * 224 * 2])
def test_quantize_dequantize(ctx, n):
    import torch
    x = orc.fill_f32((n,), 1, "q.x", -400, 400, 6)
    s = 0.0236
    dq = torch.zeros(n, dtype=torch.int8, device="cuda")
    ctx.quantize_f32_i8(_t(x), s, dq)
    ctx.sync()
    assert np.array_equal(dq.cpu().numpy(), orc.quantize(x, s))
    dx = torch.zeros(n, dtype=torch.float32, device="cuda")
    ctx.dequantize_i8_f32(dq, s, dx)
    ctx.sync()
    assert np.array_equal(dx.cpu().numpy(), orc.dequantize(dq.cpu().numpy(), s))


def test_dequantize_per_channel(ctx):
    import torch
    q = orc.fill_i8((3, 7, 5, 9), 2, "pc.q")
    s = np.linspace(0.01, 0.2, 7).astype(np.float32)
    dx = torch.zeros(q.shape, dtype=torch.float32, device="cuda")
    ctx.dequantize_i8_f32_per_channel(_t(q), _t(s), dx)
    ctx.sync()
    assert np.array_equal(dx.cpu().numpy(), orc.dequantize_per_channel(q, s))


def test_bn_relu_add_f32(ctx):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 6, 9, 7)).astype(np.float32)
    g, b, m = (rng.standard_normal(6).astype(np.float32) for _ in range(3))
    v = rng.uniform(0.5, 1.5, 6).astype(np.float32)
    dx = _t(x)
    ctx.bn_inference_f32(dx, _t(g), _t(b), _t(m), _t(v), 1e-5)
    ctx.sync()
    assert np.array_equal(dx.cpu().numpy(), orc.bn_inference_f32(x, g, b, m, v))   # same op order, IEEE div/sqrt
    ctx.relu_forward_f32(dx)
    ctx.sync()
    r = orc.relu_f32(orc.bn_inference_f32(x, g, b, m, v))
    assert np.array_equal(dx.cpu().numpy(), r)
    dy = _t(x)
    ctx.add_inplace_f32(dy, dx)
    ctx.sync()
    assert np.array_equal(dy.cpu().numpy(), orc.add_f32(x, r))


def test_relu_add_requant_i8(ctx):
    y = orc.fill_i8((1237,), 0, "ar.y")
    x = orc.fill_i8((1237,), 0, "ar.x")
    dy = _t(y)
    ctx.add_requant_i8(dy, 0.11, _t(x), 0.23, True, 0.17)
    ctx.sync()
    assert np.array_equal(dy.cpu().numpy(), orc.add_requant_i8(y, 0.11, x, 0.23, True, 0.17))
    dz = _t(y)
    ctx.relu_forward_i8(dz)
    ctx.sync()
    assert np.array_equal(dz.cpu().numpy(), np.maximum(y, 0))


@pytest.mark.parametrize("shape", [(2, 64, 112, 112), (1, 5, 7, 9), (3, 16, 8, 8)])
def test_maxpool_i8(ctx, shape):
    import torch
    x = orc.fill_i8(shape, 3, "mp.x")
    ref = orc.maxpool_i8(x)
    dy = torch.zeros(ref.shape, dtype=torch.int8, device="cuda")
    ctx.maxpool2d_3x3_s2p1_nchw_i8(_t(x), dy)
    ctx.sync()
    assert np.array_equal(dy.cpu().numpy(), ref)


def test_gap_fc_softmax(ctx):
    import torch
    x = orc.fill_i8((4, 512, 7, 7), 4, "gap.x")
    yf_ref, yq_ref = orc.gap_i8(x, 0.3, 0.05)
    dyf = torch.zeros((4, 512), dtype=torch.float32, device="cuda")
    dyq = torch.zeros((4, 512), dtype=torch.int8, device="cuda")
    ctx.gap_global_i8(_t(x), 0.3, 0.05, dyf, dyq)
    ctx.sync()
    assert np.array_equal(dyf.cpu().numpy(), yf_ref) and np.array_equal(dyq.cpu().numpy(), yq_ref)
    w = orc.fill_i8((1000, 512), 4, "fc.w", -127, 127)
    sc = np.linspace(1e-3, 2e-3, 1000).astype(np.float32)
    b = orc.fill_f32((1000,), 4, "fc.b", -64, 64, 8)
    _, lg_ref = orc.fc_i8(yq_ref, w, sc, b)
    dl = torch.zeros((4, 1000), dtype=torch.float32, device="cuda")
    ctx.fc_forward_i8(dyq, _t(w), _t(sc), _t(b), dl)
    ctx.sync()
    assert np.array_equal(dl.cpu().numpy().view(np.uint32), lg_ref.view(np.uint32))
    dp = torch.zeros((4, 1000), dtype=torch.float32, device="cuda")
    lgs = (lg_ref / np.abs(lg_ref).max() * 10).astype(np.float32)
    ctx.softmax_f32(_t(lgs), dp)
    ctx.sync()
    ref = np.stack([orc.softmax_f32(lgs[i]) for i in range(4)])
    assert np.abs(dp.cpu().numpy() - ref).max() <= 1e-6        # expf vs expf, different summation trees
    assert np.abs(dp.cpu().numpy().sum(1) - 1).max() <= 1e-5


def test_native_layout_roundtrip_and_conv(ctx):
    """dlq_act_from/to_nchw_i8 are inverse; dlq_conv2d_i8_act equals the NCHW entry point."""
    import torch
    x = orc.fill_i8((3, 128, 28, 28), 5, "nl.x")
    buf, act = ctx.new_act(3, 28, 28, 128, 1)
    ctx.act_from_nchw_i8(_t(x), act)
    back = torch.zeros(x.shape, dtype=torch.int8, device="cuda")
    ctx.act_to_nchw_i8(act, back)
    ctx.sync()
    assert np.array_equal(back.cpu().numpy(), x)
    wq = orc.fill_i8((128, 128, 3, 3), 5, "nl.w", -127, 127)
    w = ctx.pack_conv_weights_i8(wq, 1, 1)
    assert ctx.required_pad_rows(w) == 1
    alpha = np.full(128, 2.0 ** -9, np.float32)
    beta = np.zeros(128, np.float32)
    ybuf, yact = ctx.new_act(3, 28, 28, 128, 2)
    ctx.conv2d_i8_act(act, w, yact, _t(alpha), _t(beta), None, 0.0, True)
    y = torch.zeros(x.shape, dtype=torch.int8, device="cuda")
    ctx.act_to_nchw_i8(yact, y)
    ctx.sync()
    _, yref = orc.conv2d_i8(x, wq, 1, 1, alpha, beta, None, 0.0, True, want_acc=False)
    assert np.array_equal(y.cpu().numpy(), yref)
    # pad rows of the output tensor must still be zero (never written by the kernel)
    rows = ybuf[:yact.N * 0 + (2 + 3 * 30) * 28 * 128].view(2 + 3 * 30, 28 * 128)
    for n in range(3):
        assert int(rows[n * 30:n * 30 + 2].abs().sum()) == 0
    w.free()
