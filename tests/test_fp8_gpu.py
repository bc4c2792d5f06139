"""FP8 (E4M3) path, QUANT_SPEC section 6, through the C ABI vs the CPU oracle.
The tensor core accumulates E4M3 products in FP32 in an unspecified order, the oracle in double: parity is by
tolerance — raw accumulators within 2^-10 relative of the magnitude sum's scale, outputs within ONE E4M3 code on
>= 99.9 % of elements (identical on most), network logits within 1e-2 relative L2 (the figures QUANT_SPEC states)."""
import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu

SHAPES = [  # IC, H, OC, k, stride, pad, residual
    (64, 56, 64, 3, 1, 1, True),
    (64, 56, 128, 3, 2, 1, False),
    (64, 56, 128, 1, 2, 0, False),
    (128, 28, 128, 3, 1, 1, True),
    (256, 14, 512, 3, 2, 1, False),
    (512, 7, 512, 3, 1, 1, True),
    (3, 224, 64, 7, 2, 3, False),
]


def _codes(shape, seed, name, lo=-96, hi=96):
    """E4M3 codes of lattice values in [lo, hi] / 32 (exactly representable or rounded by the oracle's converter)"""
    v = orc.fill_f32(shape, seed, name, lo, hi, 5)
    return orc.quantize_e4m3(v, 1.0)


def _code_distance(a, b):
    """distance in E4M3 code steps along the value order (sign-magnitude codes -> monotone integers)"""
    def mono(c):
        c = c.astype(np.int32)
        return np.where(c & 0x80, -(c & 0x7F), c & 0x7F)
    return np.abs(mono(a) - mono(b))


@pytest.mark.parametrize("ic,h,oc,k,stride,pad,res", SHAPES)
def test_conv_fp8_matches_oracle(ic, h, oc, k, stride, pad, res):
    import torch
    import dlq_b200
    ctx = dlq_b200.Context(0)
    n = 3
    x = _codes((n, ic, h, h), 1, "fp8.x")
    wq = _codes((oc, ic, k, k), 1, "fp8.w", -64, 64)
    oh = (h + 2 * pad - k) // stride + 1
    r = _codes((n, oc, oh, oh), 1, "fp8.r") if res else None
    alpha = np.full(oc, 2.0 ** -6, np.float32) * np.linspace(0.5, 1.5, oc).astype(np.float32)
    beta = np.linspace(-3, 3, oc).astype(np.float32)
    acc_ref, y_ref = orc.conv2d_e4m3(x, wq, stride, pad, alpha, beta, r, 0.25, True)
    w = ctx.pack_conv_weights_e4m3(wq, stride, pad)
    dy = torch.empty((n, oc, oh, oh), dtype=torch.uint8, device="cuda")
    dacc = torch.empty((n, oc, oh, oh), dtype=torch.float32, device="cuda")
    ctx.conv2d_fp8(torch.from_numpy(x).cuda(), w, torch.from_numpy(alpha).cuda(), torch.from_numpy(beta).cuda(),
                   torch.from_numpy(r).cuda() if res else None, 0.25, True, dy, dacc)
    ctx.sync()
    acc = dacc.cpu().numpy()
    # fp32 accumulation of <= 4608 products of magnitude <= 9: absolute error bounded by K * 2^-24 * sum|terms|
    tol = 1e-3 * max(1.0, float(np.abs(acc_ref).max()))
    assert np.abs(acc - acc_ref).max() <= tol, f"accumulators differ by {np.abs(acc - acc_ref).max()} (tol {tol})"
    d = _code_distance(dy.cpu().numpy(), y_ref)
    assert d.max() <= 1, f"outputs differ by {d.max()} E4M3 codes"
    assert (d == 0).mean() >= 0.999, f"only {(d == 0).mean():.5f} of the outputs are identical"
    w.free()
    ctx.close()


def test_quantize_dequantize_e4m3_bit_exact():
    import torch
    import dlq_b200
    ctx = dlq_b200.Context(0)
    x = orc.fill_f32((1 << 16) + 7, 3, "fp8.q", -40000, 40000, 6)     # up to +-625: exercises saturation at 448
    x[:512] = np.linspace(-0.05, 0.05, 512, dtype=np.float32)        # subnormal range
    dq = torch.empty(x.size, dtype=torch.uint8, device="cuda")
    ctx.quantize_f32_e4m3(torch.from_numpy(x).cuda(), 0.75, dq)
    dx = torch.empty(x.size, dtype=torch.float32, device="cuda")
    ctx.dequantize_e4m3_f32(dq, 0.75, dx)
    ctx.sync()
    q_ref = orc.quantize_e4m3(x, 0.75)
    assert np.array_equal(dq.cpu().numpy(), q_ref)
    assert np.array_equal(dx.cpu().numpy().view(np.uint32), orc.dequantize_e4m3(q_ref, 0.75).view(np.uint32))
    ctx.close()


def test_resnet18_fp8_network_vs_oracle():
    import torch
    import dlq_b200
    from dlq_b200 import synth
    w = synth.make_weights(0)
    s8 = orc.fp8_act_scales(synth.load_act_scales(0))
    x = synth.make_input(0, 2)
    ref = orc.FP8Model(w, s8).forward(x, checkpoints=True)
    ctx = dlq_b200.Context(0)
    m = dlq_b200.ResNet18(ctx, w, s8, 2, fp8=True)
    dl = torch.empty((2, 1000), dtype=torch.float32, device="cuda")
    m.forward(torch.from_numpy(x).cuda(), dl)
    ctx.sync()
    for k in ("stem_pool", "layer1", "layer2", "layer3", "layer4"):
        t = torch.empty(ref[k].shape, dtype=torch.int8, device="cuda")
        m.checkpoint(k, t)
        ctx.sync()
        d = _code_distance(t.cpu().numpy().view(np.uint8), ref[k])
        # rounding differences propagate through the layers: the stated bar is on the first conv-level checkpoint
        # exactly (<= 1 code, >= 99.9 % identical) and a widening allowance afterwards
        frac_same = (d == 0).mean()
        if k == "stem_pool":
            assert d.max() <= 1 and frac_same >= 0.999, (k, d.max(), frac_same)
        else:
            assert frac_same >= 0.97 and (d <= 1).mean() >= 0.999, (k, d.max(), frac_same)
    got = dl.cpu().numpy()
    rel = np.linalg.norm(got - ref["logits"]) / np.linalg.norm(ref["logits"])
    assert rel <= 1e-2, f"logits relative L2 error {rel}"
    assert np.array_equal(got.argmax(1), ref["logits"].argmax(1))
    m.close()
    ctx.close()
