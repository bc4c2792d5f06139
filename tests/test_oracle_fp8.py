"""CPU checks of the E4M3 half of the oracle (QUANT_SPEC section 6): code table, converter, NumPy brute force."""
import numpy as np

import orc


def test_e4m3_table_and_roundtrip():
    t = orc.e4m3_table()
    assert t[0x00] == 0.0 and t[0x7E] == 448.0 and np.isnan(t[0x7F]) and t[0x01] == 2.0 ** -9 and t[0x08] == 2.0 ** -6
    assert t[0x38] == 1.0 and t[0xB8] == -1.0
    for c in list(range(0x7F)) + list(range(0x80, 0xFF)):          # every finite code converts back to itself
        assert orc.f32_to_e4m3(float(t[c])) == (c if c != 0x80 else 0x80)
    pos = t[:0x7F]
    assert np.all(np.diff(pos) > 0)                                # codes are monotone in value


def test_e4m3_rounding_is_nearest_even_and_saturating():
    t = orc.e4m3_table()
    for c in range(0x7D):
        mid = (float(t[c]) + float(t[c + 1])) / 2                  # exactly representable in fp32
        even = c if c % 2 == 0 else c + 1
        assert orc.f32_to_e4m3(mid) == even, (c, mid)
        assert orc.f32_to_e4m3(np.nextafter(np.float32(mid), np.float32(0))) == c
        assert orc.f32_to_e4m3(np.nextafter(np.float32(mid), np.float32(1e9))) == c + 1
    assert orc.f32_to_e4m3(1e9) == 0x7E and orc.f32_to_e4m3(-1e9) == 0xFE and orc.f32_to_e4m3(464.0) == 0x7E
    assert orc.f32_to_e4m3(float("nan")) & 0x7F == 0x7F


def test_conv_e4m3_against_numpy():
    rng = np.random.default_rng(0)
    t = orc.e4m3_table().astype(np.float64)
    x = rng.integers(0, 0x60, (2, 8, 9, 9), dtype=np.uint8) | (rng.integers(0, 2, (2, 8, 9, 9), dtype=np.uint8) << 7)
    w = rng.integers(0, 0x50, (16, 8, 3, 3), dtype=np.uint8) | (rng.integers(0, 2, (16, 8, 3, 3), dtype=np.uint8) << 7)
    acc, _ = orc.conv2d_e4m3(x, w, 2, 1)
    xf, wf = t[x], t[w]
    xp = np.pad(xf, ((0, 0), (0, 0), (1, 1), (1, 1)))
    ref = np.zeros((2, 16, 5, 5))
    for oh in range(5):
        for ow in range(5):
            patch = xp[:, :, 2 * oh:2 * oh + 3, 2 * ow:2 * ow + 3]
            ref[:, :, oh, ow] = np.einsum("nchw,ochw->no", patch, wf)
    assert np.array_equal(acc, ref.astype(np.float32))


def test_fp8_network_close_to_fp32():
    """E4M3 network vs the FP32 oracle on the synthetic model: logits cosine (the dequantise + tolerance link)."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from dlq_b200 import synth
    w = synth.make_weights(0, fill=orc.fill_f32)
    x = synth.make_input(0, 1, fill=orc.fill_f32)
    s8 = orc.fp8_act_scales(synth.load_act_scales(0))
    a = orc.FP8Model(w, s8).forward(x)["logits"][0].astype(np.float64)
    b = orc.F32Model(w).forward(x)["logits"][0].astype(np.float64)
    cos = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))
    assert cos > 0.97, cos
