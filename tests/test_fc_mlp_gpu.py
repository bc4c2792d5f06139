"""GPU: fully-connected layers on the tensor-core GEMM core (dlq_fc_forward_i8_tc / _fp8: the conv kernel as a 1x1
convolution) and the MNIST MLP forward behind the C ABI (dlq_mlp_*), SURVEY 8f-4.
Oracle: the restated INT8 FC / conv epilogue (orc.fc_i8, orc.conv2d_i8 - QUANT_SPEC 3, 5), bit-exact; the reference's own
CPU forward (MN/v3.c forward_timed, golden tests/golden/mnist_v3_seed.npz) by tolerance (FP32 vs INT8)."""
import os

import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("n,i,o", [(1, 512, 1000), (5, 512, 1000), (300, 784, 256), (129, 100, 10), (1024, 256, 10),
                                   (7, 64, 64), (260, 2048, 128)])
def test_fc_tc_bit_exact_vs_oracle_and_dp4a(ctx, n, i, o):
    """tcgen05 FC == CPU oracle == the dp4a operator (dlq_fc_forward_i8), bit for bit, padded and unpadded shapes"""
    import torch
    g = orc.fill_i8((n, i), 4, "fc.g")
    wq = orc.fill_i8((o, i), 4, "fc.w", -127, 127)
    sc = (np.linspace(0.5, 1.5, o) * 2.0 ** -9).astype(np.float32)
    bias = np.linspace(-3, 3, o).astype(np.float32)
    acc_ref, ref = orc.fc_i8(g, wq, sc, bias)
    w = ctx.pack_fc_weights_i8(wq)
    dg, dsc, db = torch.from_numpy(g).cuda(), torch.from_numpy(sc).cuda(), torch.from_numpy(bias).cuda()
    got = torch.full((n, o), float("nan"), dtype=torch.float32, device="cuda")
    ctx.fc_forward_i8_tc(dg, w, dsc, db, got)
    ctx.sync()
    assert np.array_equal(got.cpu().numpy().view(np.uint32), ref.view(np.uint32))
    old = torch.empty((n, o), dtype=torch.float32, device="cuda")
    ctx.fc_forward_i8(dg, torch.from_numpy(wq).cuda(), dsc, db, old)
    ctx.sync()
    assert torch.equal(got.view(torch.int32), old.view(torch.int32))
    w.free()


def test_fc_fp8_vs_oracle(ctx):
    """E4M3 FC on kind::f8f6f4: FP32 accumulation in unspecified order vs the oracle's double - tolerance (QUANT_SPEC 7)"""
    import torch
    n, i, o = 64, 512, 1000
    g = orc.quantize_e4m3(orc.fill_f32((n, i), 2, "fc8.g", -96, 96, 5), 1.0)
    wq = orc.quantize_e4m3(orc.fill_f32((o, i), 2, "fc8.w", -64, 64, 5), 1.0)
    sc = np.full(o, 2.0 ** -6, np.float32)
    bias = np.linspace(-1, 1, o).astype(np.float32)
    gf = orc.dequantize_e4m3(g, 1.0).astype(np.float64)
    wf = orc.dequantize_e4m3(wq, 1.0).astype(np.float64)
    ref = (gf @ wf.T) * sc.astype(np.float64) + bias.astype(np.float64)
    w = ctx.pack_fc_weights_e4m3(wq)
    got = torch.empty((n, o), dtype=torch.float32, device="cuda")
    ctx.fc_forward_fp8(torch.from_numpy(g).cuda(), w, torch.from_numpy(sc).cuda(), torch.from_numpy(bias).cuda(), got)
    ctx.sync()
    err = np.abs(got.cpu().numpy().astype(np.float64) - ref).max()
    assert err <= 1e-3 * max(1.0, np.abs(ref).max()), err
    # run-to-run determinism (one issuer per accumulator)
    again = torch.empty_like(got)
    ctx.fc_forward_fp8(torch.from_numpy(g).cuda(), w, torch.from_numpy(sc).cuda(), torch.from_numpy(bias).cuda(), again)
    ctx.sync()
    assert torch.equal(got.view(torch.int32), again.view(torch.int32))
    w.free()


def test_workspace_reserve_contract():
    """SURVEY 8b "Ownership": after dlq_workspace_reserve the per-layer entry points never allocate - the workspace
    keeps its size and address - and a call that needs more fails with code 1 instead of allocating"""
    import torch
    import dlq_b200
    ctx = dlq_b200.Context(0)
    wq = orc.fill_i8((64, 64, 3, 3), 0, "ws.w", -127, 127)
    w = ctx.pack_conv_weights_i8(wq, 1, 1)
    need = ctx.conv2d_workspace_bytes(w, 4, 56, 56, has_residual=True, want_acc=True)
    assert need > 4 * 64 * 56 * 56 * 6
    ctx.workspace_reserve(need)
    assert ctx.workspace_bytes == need
    x = orc.fill_i8((4, 64, 56, 56), 0, "ws.x")
    r = orc.fill_i8((4, 64, 56, 56), 0, "ws.r")
    alpha = np.full(64, 2.0 ** -8, np.float32)
    beta = np.zeros(64, np.float32)
    dy = torch.empty((4, 64, 56, 56), dtype=torch.int8, device="cuda")
    dacc = torch.empty((4, 64, 56, 56), dtype=torch.int32, device="cuda")
    ctx.conv2d_i8(torch.from_numpy(x).cuda(), w, torch.from_numpy(alpha).cuda(), torch.from_numpy(beta).cuda(),
                  torch.from_numpy(r).cuda(), 0.5, True, dy, dacc)
    ctx.sync()
    acc_ref, y_ref = orc.conv2d_i8(x, wq, 1, 1, alpha, beta, r, 0.5, True)
    assert np.array_equal(dacc.cpu().numpy(), acc_ref) and np.array_equal(dy.cpu().numpy(), y_ref)
    assert ctx.workspace_bytes == need
    x8 = torch.zeros((32, 64, 56, 56), dtype=torch.int8, device="cuda")
    with pytest.raises(dlq_b200.DlqError) as ei:
        ctx.conv2d_i8(x8, w, torch.from_numpy(alpha).cuda(), torch.from_numpy(beta).cuda(), None, 0.0, True,
                      torch.empty((32, 64, 56, 56), dtype=torch.int8, device="cuda"), None)
    assert ei.value.code == 1 and "workspace" in str(ei.value)
    w.free()
    ctx.close()


def _mnist_oracle(m, x, w1, b1, w2, b2):
    """the arithmetic include/dlq.h states for dlq_mlp_forward, restated with the oracle's operators"""
    s_w1, s_w2 = m.weight_scales(1), m.weight_scales(2)
    w1q = np.clip(np.rint(w1.T * (1.0 / s_w1.astype(np.float64)).astype(np.float32)[:, None]), -127, 127).astype(np.int8)
    w2q = np.clip(np.rint(w2.T * (1.0 / s_w2.astype(np.float64)).astype(np.float32)[:, None]), -127, 127).astype(np.int8)
    xq = orc.quantize(x, float(m.s_x))
    alpha = (np.float64(m.s_x) * s_w1.astype(np.float64) / np.float64(m.s_h)).astype(np.float32)
    beta = (b1.astype(np.float64) / np.float64(m.s_h)).astype(np.float32)
    # FC1 = 1x1 conv with the QUANT_SPEC 3 epilogue (bias + ReLU + requantisation in one fmaf)
    _, hq = orc.conv2d_i8(xq[:, :, None, None], w1q[:, :, None, None], 1, 0, alpha, beta, None, 0.0, True)
    hq = hq[:, :, 0, 0]
    sc2 = (np.float64(m.s_h) * s_w2.astype(np.float64)).astype(np.float32)
    _, z = orc.fc_i8(hq, w2q, sc2, b2)
    return xq, hq, z


def test_mnist_mlp_c_entry_vs_oracle_and_reference():
    import torch
    import dlq_b200
    from dlq_b200.mnist import MnistMLP
    g = np.load(os.path.join(GOLD, "mnist_v3_seed.npz"))
    x, w1, b1, w2, b2 = (g[k] for k in ("x", "w1", "b1", "w2", "b2"))
    B = x.shape[0]
    ctx = dlq_b200.Context(0)
    m = MnistMLP(ctx, w1, b1, w2, b2, x, max_batch=B)
    dz = torch.empty((B, 10), dtype=torch.float32, device="cuda")
    dp = torch.empty((B, 10), dtype=torch.float32, device="cuda")
    m.forward(torch.from_numpy(x).cuda(), dz, dp)
    dxq = torch.empty((B, 784), dtype=torch.uint8, device="cuda")
    dhq = torch.empty((B, 256), dtype=torch.uint8, device="cuda")
    m.checkpoint("input", dxq)
    m.checkpoint("hidden", dhq)
    ctx.sync()
    z, p = dz.cpu().numpy(), dp.cpu().numpy()
    # (a) oracle restatement: quantised input, hidden activations and logits bit-exact
    xq, hq, z_ref = _mnist_oracle(m, x, w1, b1, w2, b2)
    assert np.array_equal(dxq.cpu().numpy().view(np.int8), xq)
    assert np.array_equal(dhq.cpu().numpy().view(np.int8), hq)
    assert np.array_equal(z.view(np.uint32), z_ref.view(np.uint32))
    assert np.allclose(p, np.stack([orc.softmax_f32(r) for r in z_ref]), atol=2e-6)
    # a smaller, ragged batch through the same model
    m.forward(torch.from_numpy(x[:77]).cuda(), dz[:77], None)
    ctx.sync()
    assert np.array_equal(dz[:77].cpu().numpy().view(np.uint32), z_ref[:77].view(np.uint32))
    # (b) the reference's FP32 CPU forward (v3.c stores probabilities, not logits)
    ref_p = g["out"]
    assert np.array_equal(p.argmax(1), ref_p.argmax(1))
    assert np.abs(p - ref_p).max() < 0.05
    m.close()
    # (c) E4M3 variant: tolerance vs the reference probabilities, deterministic run to run
    m8 = MnistMLP(ctx, w1, b1, w2, b2, x, max_batch=B, fp8=True)
    p8 = torch.empty((B, 10), dtype=torch.float32, device="cuda")
    p8b = torch.empty((B, 10), dtype=torch.float32, device="cuda")
    m8.forward(torch.from_numpy(x).cuda(), None, p8)
    m8.forward(torch.from_numpy(x).cuda(), None, p8b)
    ctx.sync()
    assert torch.equal(p8.view(torch.int32), p8b.view(torch.int32))
    # (3 mantissa bits against near-tied logits of a random-init MLP: agreement, not identity)
    assert (p8.cpu().numpy().argmax(1) == ref_p.argmax(1)).mean() >= 0.90
    assert np.abs(p8.cpu().numpy() - ref_p).max() < 0.25
    m8.close()
    ctx.close()
