"""CPU: the INT8 half of the oracle against the written spec (spec/QUANT_SPEC.md), NumPy brute force and the
committed golden vector.  (Parity unpinned by the reference: it has no quantised code.)"""
import hashlib
import json
import os

import numpy as np

import orc
from dlq_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_quantize_round_half_even_and_clamp():
    s = np.float32(0.5)
    x = np.array([0.25, 0.75, 1.25, -0.25, -0.75, 1000.0, -1000.0, 63.49, 63.75], np.float32)
    q = orc.quantize(x, s)
    assert q.tolist() == [0, 2, 2, 0, -2, 127, -128, 127, 127]     # ties to even, saturating
    assert np.array_equal(orc.dequantize(q, s), q.astype(np.float32) * s)
    assert orc.quantize(np.zeros(0, np.float32), s).size == 0     # empty input


def test_weight_quantisation_per_channel():
    rng = np.random.default_rng(1)
    w = rng.standard_normal((8, 5, 3, 3)).astype(np.float32)
    w[3] = 0                                                         # all-zero row -> scale 1, q 0
    q, s = orc.quantize_weights(w)
    am = np.abs(w.reshape(8, -1)).max(1)
    assert np.array_equal(s[am > 0], (am[am > 0].astype(np.float64) / 127).astype(np.float32))
    assert s[3] == 1.0 and not q[3].any()
    assert np.abs(q).max() == 127 and q.min() >= -127
    deq = orc.dequantize_per_channel(q.reshape(1, 8, -1), s).reshape(w.shape)   # [N=1, C=OC, HW=K]
    assert np.abs(deq - w).max() <= s[am > 0].max() / 2 + 1e-7


def test_conv_i8_vs_numpy_bruteforce():
    rng = np.random.default_rng(2)
    for (n, c, h, w_, oc, k, s, p) in [(2, 4, 7, 6, 5, 3, 1, 1), (1, 3, 9, 9, 4, 7, 2, 3), (2, 6, 8, 8, 3, 1, 2, 0)]:
        x = rng.integers(-128, 128, (n, c, h, w_), dtype=np.int8)
        wq = rng.integers(-127, 128, (oc, c, k, k), dtype=np.int8)
        alpha = rng.uniform(1e-3, 3e-3, oc).astype(np.float32)
        beta = rng.uniform(-20, 20, oc).astype(np.float32)
        oh, ow = (h + 2 * p - k) // s + 1, (w_ + 2 * p - k) // s + 1
        res = rng.integers(-128, 128, (n, oc, oh, ow), dtype=np.int8)
        acc, y = orc.conv2d_i8(x, wq, s, p, alpha, beta, res, 0.3, True)
        xp = np.pad(x.astype(np.int64), ((0, 0), (0, 0), (p, p), (p, p)))
        ref = np.zeros((n, oc, oh, ow), np.int64)
        for i in range(oh):
            for j in range(ow):
                patch = xp[:, :, i * s:i * s + k, j * s:j * s + k]
                ref[:, :, i, j] = np.einsum("nckl,ockl->no", patch, wq.astype(np.int64))
        assert np.array_equal(acc, ref.astype(np.int32))
        # epilogue per spec §3, in float32 with explicit fma via float64 (exact for these magnitudes)
        t = (ref.astype(np.float64) * alpha.astype(np.float64)[None, :, None, None] + beta.astype(np.float64)[None, :, None, None]).astype(np.float32)
        t = (res.astype(np.float64) * np.float64(np.float32(0.3)) + t.astype(np.float64)).astype(np.float32)
        yref = np.clip(np.rint(t), 0, 127).astype(np.int8)
        assert np.array_equal(y, yref)


def test_maxpool_gap_fc_add_i8():
    rng = np.random.default_rng(3)
    x = rng.integers(-128, 128, (2, 5, 9, 7), dtype=np.int8)
    mp = orc.maxpool_i8(x)
    assert np.array_equal(mp, orc.maxpool_f32(x.astype(np.float32)).astype(np.int8))   # monotone => exact
    yf, yq = orc.gap_i8(x, 0.05, 0.02)
    s = np.float32(np.float64(np.float32(0.05)) / 63)
    assert np.array_equal(yf, x.sum((2, 3)).astype(np.float32) * s)
    g = rng.integers(-128, 128, (3, 16), dtype=np.int8)
    w = rng.integers(-127, 128, (10, 16), dtype=np.int8)
    sc, b = rng.uniform(1e-3, 2e-3, 10).astype(np.float32), rng.standard_normal(10).astype(np.float32)
    acc, lg = orc.fc_i8(g, w, sc, b)
    assert np.array_equal(acc, g.astype(np.int32) @ w.astype(np.int32).T)
    y = rng.integers(-128, 128, 100, dtype=np.int8)
    x2 = rng.integers(-128, 128, 100, dtype=np.int8)
    out = orc.add_requant_i8(y, 0.1, x2, 0.2, True, 0.15)
    assert out.min() >= 0 and out.max() <= 127


def test_int8_network_tracks_fp32_and_golden():
    w = synth.make_weights(0, fill=orc.fill_f32)
    scales = synth.load_act_scales(0)
    x = synth.make_input(0, 2, fill=orc.fill_f32)
    out = orc.I8Model(w, scales).forward(x, checkpoints=True)
    g = np.load(os.path.join(GOLD, "i8_seed0_n2.npz"))
    assert np.array_equal(out["logits"].view(np.uint32), g["logits"].view(np.uint32))
    dig = json.loads(str(g["digests"]))
    for k, d in dig.items():
        assert hashlib.sha256(out[k].tobytes()).hexdigest() == d, k
    f = orc.F32Model(w).forward(x, checkpoints=True)
    # dequantised int8 checkpoints vs fp32: SQNR bound; logits: cosine
    s_l4 = scales[orc.ACT_BLOCK0 + 3 * 7 + 2]
    deq = out["layer4"].astype(np.float32) * s_l4
    sqnr = 10 * np.log10((f["layer4"] ** 2).sum() / ((f["layer4"] - deq) ** 2).sum())
    assert sqnr > 15, sqnr
    for i in range(2):
        cos = float(np.dot(f["logits"][i], out["logits"][i]) / (np.linalg.norm(f["logits"][i]) * np.linalg.norm(out["logits"][i])))
        assert cos > 0.98, cos
