"""ctypes binding of the CPU oracle (oracle/_build/libdlq_oracle.so).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "_build", "libdlq_oracle.so")

NUM_CONVS, NUM_BLOCKS, NUM_ACTS = 25, 8, 27
ACT_INPUT, ACT_STEM, ACT_BLOCK0, ACT_GAP = 0, 1, 2, 26


class ConvBN(C.Structure):
    _fields_ = [("ic", C.c_int), ("oc", C.c_int), ("k", C.c_int), ("stride", C.c_int), ("pad", C.c_int),
                ("w", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("mean", C.c_void_p),
                ("var", C.c_void_p)]


class ResNetF32(C.Structure):
    _fields_ = [("convs", ConvBN * NUM_CONVS), ("fc_w", C.c_void_p), ("fc_b", C.c_void_p)]


class CkF32(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("stem_pool", "layer1", "layer2", "layer3", "layer4", "gap", "absmax")]


class Epilogue(C.Structure):
    _fields_ = [("alpha", C.c_void_p), ("beta", C.c_void_p), ("residual", C.c_void_p), ("res_mul", C.c_float),
                ("relu", C.c_int)]


class ConvQ(C.Structure):
    _fields_ = [("ic", C.c_int), ("oc", C.c_int), ("k", C.c_int), ("stride", C.c_int), ("pad", C.c_int),
                ("w", C.c_void_p), ("alpha", C.c_void_p), ("beta", C.c_void_p)]


class ResNetI8(C.Structure):
    _fields_ = [("convs", ConvQ * NUM_CONVS), ("act_scale", C.c_float * NUM_ACTS), ("fc_w", C.c_void_p),
                ("fc_scale", C.c_void_p), ("fc_b", C.c_void_p)]


class CkI8(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("stem_pool", "layer1", "layer2", "layer3", "layer4", "gap")]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.orc_inv_scale.restype = C.c_float
        _lib.orc_inv_scale.argtypes = [C.c_float]
        _lib.orc_num_threads.restype = C.c_int
        _lib.orc_res_mul.restype = C.c_float
        _lib.orc_res_mul.argtypes = [C.c_float, C.c_float]
        _lib.orc_f32_to_e4m3.restype = C.c_uint8
        _lib.orc_f32_to_e4m3.argtypes = [C.c_float]
        _lib.orc_e4m3_to_f32.restype = C.c_float
        _lib.orc_e4m3_to_f32.argtypes = [C.c_uint8]
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def fill_f32(shape, seed, name, lo, hi, shift) -> np.ndarray:
    a = np.empty(shape, dtype=np.float32)
    lib().orc_fill_lattice_f32(_p(a), C.c_size_t(a.size), C.c_uint64(seed), name.encode(), lo, hi, shift)
    return a


def fill_i8(shape, seed, name, lo=-128, hi=127) -> np.ndarray:
    a = np.empty(shape, dtype=np.int8)
    lib().orc_fill_lattice_i8(_p(a), C.c_size_t(a.size), C.c_uint64(seed), name.encode(), lo, hi)
    return a


def inv_scale(s: float) -> float:
    return float(lib().orc_inv_scale(C.c_float(s)))


# ------------------------------------------------------------------ FP32 ops
def conv2d_f32(x, w, stride, pad):
    x, w = f32(x), f32(w)
    n, c, h, ww = x.shape
    oc, _, kh, kw = w.shape
    oh, ow = (h + 2 * pad - kh) // stride + 1, (ww + 2 * pad - kw) // stride + 1
    y = np.empty((n, oc, oh, ow), dtype=np.float32)
    lib().orc_conv2d_f32(_p(x), n, c, h, ww, _p(w), oc, kh, kw, stride, stride, pad, pad, _p(y))
    return y


def sgemm_f32(a, b):
    a, b = f32(a), f32(b)
    m, k = a.shape
    n = b.shape[1]
    c = np.empty((m, n), dtype=np.float32)
    lib().orc_sgemm_f32(_p(a), _p(b), _p(c), m, n, k)
    return c


def bn_inference_f32(x, g, b, m, v, eps=1e-5):
    x = f32(x).copy()
    n, c, oh, ow = x.shape
    lib().orc_bn_inference_f32(_p(x), _p(f32(g)), _p(f32(b)), _p(f32(m)), _p(f32(v)), C.c_float(eps), n, c, oh, ow)
    return x


def relu_f32(x):
    x = f32(x).copy()
    lib().orc_relu_f32(_p(x), C.c_size_t(x.size))
    return x


def add_f32(y, x):
    y = f32(y).copy()
    lib().orc_add_f32(_p(y), _p(f32(x)), C.c_size_t(y.size))
    return y


def maxpool_f32(x):
    x = f32(x)
    n, c, h, w = x.shape
    oh, ow = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
    y = np.empty((n, c, oh, ow), dtype=np.float32)
    lib().orc_maxpool3x3s2p1_f32(_p(x), n, c, h, w, _p(y))
    return y


def gap_f32(x):
    x = f32(x)
    n, c, h, w = x.shape
    y = np.empty((n, c), dtype=np.float32)
    lib().orc_gap_f32(_p(x), n, c, h, w, _p(y))
    return y


def fc_f32(gap, w, b):
    gap, w, b = f32(gap), f32(w), f32(b)
    n, i = gap.shape
    o = w.shape[0]
    out = np.empty((n, o), dtype=np.float32)
    lib().orc_fc_f32(_p(gap), _p(w), _p(b), _p(out), n, o, i)
    return out


def softmax_f32(x):
    x = f32(x)
    y = np.empty_like(x)
    lib().orc_softmax_f32(_p(x), x.size, _p(y))
    return y


def mnist_forward(x, w1, b1, w2, b2):
    x, w1, b1, w2, b2 = map(f32, (x, w1, b1, w2, b2))
    batch, ind = x.shape
    hid, outd = w1.shape[1], w2.shape[1]
    hidden = np.empty((batch, hid), dtype=np.float32)
    out = np.empty((batch, outd), dtype=np.float32)
    lib().orc_mnist_mlp_forward(_p(x), _p(w1), _p(b1), _p(w2), _p(b2), batch, ind, hid, outd, _p(hidden), _p(out))
    return hidden, out


# ------------------------------------------------------------------ INT8 ops
def quantize(x, scale, lo=-128, hi=127):
    x = f32(x)
    q = np.empty(x.shape, dtype=np.int8)
    lib().orc_quantize_f32_i8(_p(x), C.c_size_t(x.size), C.c_float(inv_scale(scale)), lo, hi, _p(q))
    return q


def dequantize(q, scale):
    q = np.ascontiguousarray(q, dtype=np.int8)
    x = np.empty(q.shape, dtype=np.float32)
    lib().orc_dequantize_i8_f32(_p(q), C.c_size_t(q.size), C.c_float(scale), _p(x))
    return x


def dequantize_per_channel(q, scales):
    q = np.ascontiguousarray(q, dtype=np.int8)
    n, c = q.shape[:2]
    hw = q.size // max(1, n * c)
    x = np.empty(q.shape, dtype=np.float32)
    lib().orc_dequantize_i8_f32_per_channel(_p(q), n, c, hw, _p(f32(scales)), _p(x))
    return x


def quantize_weights(w):
    w = f32(w)
    oc = w.shape[0]
    k = w.size // oc
    q = np.empty(w.shape, dtype=np.int8)
    s = np.empty(oc, dtype=np.float32)
    lib().orc_quantize_weights_per_channel(_p(w), oc, k, _p(q), _p(s))
    return q, s


def fold_bn(g, b, m, v, s_w, s_x, s_y, eps=1e-5):
    oc = len(s_w)
    alpha = np.empty(oc, dtype=np.float32)
    beta = np.empty(oc, dtype=np.float32)
    lib().orc_fold_bn(_p(f32(g)), _p(f32(b)), _p(f32(m)), _p(f32(v)), C.c_float(eps), _p(f32(s_w)), C.c_float(s_x),
                      C.c_float(s_y), oc, _p(alpha), _p(beta))
    return alpha, beta


def res_mul(s_r, s_y) -> float:
    return float(lib().orc_res_mul(C.c_float(s_r), C.c_float(s_y)))


def conv2d_i8(x, wq, stride, pad, alpha=None, beta=None, residual=None, res_mul=0.0, relu=False, want_acc=True):
    """returns (acc int32 NCHW or None, y int8 NCHW or None)"""
    x = np.ascontiguousarray(x, dtype=np.int8)
    wq = np.ascontiguousarray(wq, dtype=np.int8)
    n, c, h, w = x.shape
    oc, _, kh, kw = wq.shape
    oh, ow = (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kw) // stride + 1
    acc = np.empty((n, oc, oh, ow), dtype=np.int32) if want_acc else None
    y, ep = None, None
    keep = []
    if alpha is not None:
        y = np.empty((n, oc, oh, ow), dtype=np.int8)
        a, b = f32(alpha), f32(beta)
        r = None if residual is None else np.ascontiguousarray(residual, dtype=np.int8)
        keep = [a, b, r]
        ep = Epilogue(_p(a), _p(b), _p(r), res_mul, int(relu))
    lib().orc_conv2d_i8(_p(x), n, c, h, w, _p(wq), oc, kh, kw, stride, stride, pad, pad,
                        C.byref(ep) if ep is not None else None, _p(acc), _p(y))
    return acc, y


# ------------------------------------------------------------------ FP8 (E4M3) ops, QUANT_SPEC section 6
def f32_to_e4m3(v: float) -> int:
    return int(lib().orc_f32_to_e4m3(C.c_float(v)))


def e4m3_table() -> np.ndarray:
    """float value of each of the 256 E4M3 codes"""
    return np.array([lib().orc_e4m3_to_f32(C.c_uint8(i)) for i in range(256)], dtype=np.float32)


def quantize_e4m3(x, scale):
    x = f32(x)
    q = np.empty(x.shape, dtype=np.uint8)
    lib().orc_quantize_f32_e4m3(_p(x), C.c_size_t(x.size), C.c_float(inv_scale(scale)), _p(q))
    return q


def dequantize_e4m3(q, scale):
    q = np.ascontiguousarray(q, dtype=np.uint8)
    x = np.empty(q.shape, dtype=np.float32)
    lib().orc_dequantize_e4m3_f32(_p(q), C.c_size_t(q.size), C.c_float(scale), _p(x))
    return x


def quantize_weights_e4m3(w):
    w = f32(w)
    oc = w.shape[0]
    k = w.size // oc
    q = np.empty(w.shape, dtype=np.uint8)
    s = np.empty(oc, dtype=np.float32)
    lib().orc_quantize_weights_per_channel_e4m3(_p(w), oc, k, _p(q), _p(s))
    return q, s


def conv2d_e4m3(x, wq, stride, pad, alpha=None, beta=None, residual=None, res_mul=0.0, relu=False, want_acc=True):
    """x, wq, residual: E4M3 codes (uint8); returns (acc float32 NCHW or None, y uint8 NCHW or None)"""
    x = np.ascontiguousarray(x, dtype=np.uint8)
    wq = np.ascontiguousarray(wq, dtype=np.uint8)
    n, c, h, w = x.shape
    oc, _, kh, kw = wq.shape
    oh, ow = (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kw) // stride + 1
    acc = np.empty((n, oc, oh, ow), dtype=np.float32) if want_acc else None
    y, ep = None, None
    keep = []
    if alpha is not None:
        y = np.empty((n, oc, oh, ow), dtype=np.uint8)
        a, b = f32(alpha), f32(beta)
        r = None if residual is None else np.ascontiguousarray(residual, dtype=np.uint8)
        keep = [a, b, r]
        ep = Epilogue(_p(a), _p(b), _p(r), res_mul, int(relu))
    lib().orc_conv2d_e4m3(_p(x), n, c, h, w, _p(wq), oc, kh, kw, stride, stride, pad, pad,
                          C.byref(ep) if ep is not None else None, _p(acc), _p(y))
    return acc, y


def add_requant_i8(y, y_scale, x, x_scale, relu, out_scale):
    y = np.ascontiguousarray(y, dtype=np.int8).copy()
    x = np.ascontiguousarray(x, dtype=np.int8)
    lib().orc_add_requant_i8(_p(y), C.c_float(y_scale), _p(x), C.c_float(x_scale), C.c_size_t(y.size), int(relu),
                             C.c_float(out_scale))
    return y


def maxpool_i8(x):
    x = np.ascontiguousarray(x, dtype=np.int8)
    n, c, h, w = x.shape
    oh, ow = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
    y = np.empty((n, c, oh, ow), dtype=np.int8)
    lib().orc_maxpool3x3s2p1_i8(_p(x), n, c, h, w, _p(y))
    return y


def gap_i8(x, in_scale, out_scale):
    x = np.ascontiguousarray(x, dtype=np.int8)
    n, c, h, w = x.shape
    yf = np.empty((n, c), dtype=np.float32)
    yq = np.empty((n, c), dtype=np.int8)
    s_over_hw = np.float32(np.float64(np.float32(in_scale)) / np.float64(h * w))
    lib().orc_gap_i8(_p(x), n, c, h, w, C.c_float(s_over_hw), C.c_float(inv_scale(out_scale)), _p(yf), _p(yq))
    return yf, yq


def fc_i8(g, w, scale, bias):
    g = np.ascontiguousarray(g, dtype=np.int8)
    w = np.ascontiguousarray(w, dtype=np.int8)
    n, i = g.shape
    o = w.shape[0]
    acc = np.empty((n, o), dtype=np.int32)
    logits = np.empty((n, o), dtype=np.float32)
    lib().orc_fc_i8(_p(g), _p(w), _p(f32(scale)), _p(f32(bias)), n, o, i, _p(acc), _p(logits))
    return acc, logits


# ------------------------------------------------------------------ whole network
def _geometry():
    import sys
    sys.path.insert(0, ROOT)
    from dlq_b200.synth import conv_geometry, conv_keys
    return conv_geometry(), conv_keys()


class F32Model:
    def __init__(self, weights: Dict[str, np.ndarray]):
        geo, keys = _geometry()
        self.keep = []
        self.s = ResNetF32()
        for idx in range(NUM_CONVS):
            cb = self.s.convs[idx]
            if idx not in geo:
                cb.w = None
                continue
            ic, oc, k, st, p = geo[idx]
            wk, bn = keys[idx]
            arrs = [f32(weights[wk])] + [f32(weights[bn + s]) for s in (".weight", ".bias", ".running_mean", ".running_var")]
            self.keep += arrs
            cb.ic, cb.oc, cb.k, cb.stride, cb.pad = ic, oc, k, st, p
            cb.w, cb.gamma, cb.beta, cb.mean, cb.var = [a.ctypes.data for a in arrs]
        fw, fb = f32(weights["fc.weight"]), f32(weights["fc.bias"])
        self.keep += [fw, fb]
        self.s.fc_w, self.s.fc_b = fw.ctypes.data, fb.ctypes.data

    def forward(self, x, checkpoints=False, absmax=None):
        x = f32(x)
        n = x.shape[0]
        logits = np.empty((n, 1000), dtype=np.float32)
        ck = CkF32()
        out = {}
        if checkpoints:
            shapes = {"stem_pool": (n, 64, 56, 56), "layer1": (n, 64, 56, 56), "layer2": (n, 128, 28, 28),
                      "layer3": (n, 256, 14, 14), "layer4": (n, 512, 7, 7), "gap": (n, 512)}
            for k, shp in shapes.items():
                out[k] = np.empty(shp, dtype=np.float32)
                setattr(ck, k, out[k].ctypes.data)
        if absmax is not None:
            ck.absmax = absmax.ctypes.data
        lib().orc_resnet18_f32_forward(C.byref(self.s), _p(x), n, _p(logits), C.byref(ck))
        out["logits"] = logits
        return out


def calibrate(weights, x_calib) -> np.ndarray:
    """activation scales = absmax / 127 per tensor over the calibration batch (FP32 oracle)"""
    m = F32Model(weights)
    am = np.zeros(NUM_ACTS, dtype=np.float32)
    for i in range(0, x_calib.shape[0], 4):
        m.forward(x_calib[i:i + 4], absmax=am)
    am = np.where(am > 0, am, np.float32(1.0)).astype(np.float32)
    return (am.astype(np.float64) / 127.0).astype(np.float32)


def fp8_act_scales(int8_scales) -> np.ndarray:
    """E4M3 activation scales from the INT8 calibration: the same absmax, mapped to 448 instead of 127"""
    return (np.asarray(int8_scales, dtype=np.float64) * 127.0 / 448.0).astype(np.float32)


class I8Model:
    """INT8 oracle model: quantised weights + folded constants from fp32 weights and activation scales."""
    QUANT_W = staticmethod(lambda w: quantize_weights(w))
    FORWARD = "orc_resnet18_i8_forward"
    CK_DTYPE = np.int8

    def __init__(self, weights: Dict[str, np.ndarray], act_scale):
        geo, keys = _geometry()
        self.keep = []
        self.s = ResNetI8()
        S = np.asarray(act_scale, dtype=np.float32)
        for i in range(NUM_ACTS):
            self.s.act_scale[i] = float(S[i])
        s_in = {0: S[ACT_INPUT]}
        s_out = {0: S[ACT_STEM]}
        s_cur = S[ACT_STEM]
        for b in range(NUM_BLOCKS):
            s_in[1 + 3 * b], s_out[1 + 3 * b] = s_cur, S[ACT_BLOCK0 + 3 * b]
            s_in[2 + 3 * b], s_out[2 + 3 * b] = S[ACT_BLOCK0 + 3 * b], S[ACT_BLOCK0 + 3 * b + 2]
            s_in[3 + 3 * b], s_out[3 + 3 * b] = s_cur, S[ACT_BLOCK0 + 3 * b + 1]
            s_cur = S[ACT_BLOCK0 + 3 * b + 2]
        self.wq, self.alpha, self.beta, self.s_w = {}, {}, {}, {}
        for idx in range(NUM_CONVS):
            cq = self.s.convs[idx]
            if idx not in geo:
                cq.w = None
                continue
            ic, oc, k, st, p = geo[idx]
            wk, bn = keys[idx]
            q, sw = self.QUANT_W(weights[wk])
            a, bt = fold_bn(weights[bn + ".weight"], weights[bn + ".bias"], weights[bn + ".running_mean"],
                            weights[bn + ".running_var"], sw, float(s_in[idx]), float(s_out[idx]))
            self.wq[idx], self.alpha[idx], self.beta[idx], self.s_w[idx] = q, a, bt, sw
            cq.ic, cq.oc, cq.k, cq.stride, cq.pad = ic, oc, k, st, p
            cq.w, cq.alpha, cq.beta = q.ctypes.data, a.ctypes.data, bt.ctypes.data
        self.fc_q, fc_sw = self.QUANT_W(weights["fc.weight"])
        self.fc_scale = (np.float64(S[ACT_GAP]) * fc_sw.astype(np.float64)).astype(np.float32)
        self.fc_b = f32(weights["fc.bias"])
        self.s.fc_w, self.s.fc_scale, self.s.fc_b = self.fc_q.ctypes.data, self.fc_scale.ctypes.data, self.fc_b.ctypes.data

    def forward(self, x, checkpoints=False):
        x = f32(x)
        n = x.shape[0]
        logits = np.empty((n, 1000), dtype=np.float32)
        ck = CkI8()
        out = {}
        if checkpoints:
            shapes = {"stem_pool": (n, 64, 56, 56), "layer1": (n, 64, 56, 56), "layer2": (n, 128, 28, 28),
                      "layer3": (n, 256, 14, 14), "layer4": (n, 512, 7, 7), "gap": (n, 512)}
            for k, shp in shapes.items():
                out[k] = np.empty(shp, dtype=self.CK_DTYPE)
                setattr(ck, k, out[k].ctypes.data)
        getattr(lib(), self.FORWARD)(C.byref(self.s), _p(x), n, _p(logits), C.byref(ck))
        out["logits"] = logits
        return out


class FP8Model(I8Model):
    """E4M3 oracle model (QUANT_SPEC section 6): same wiring, E4M3 codes, double-accumulated products.
    act_scale: per-tensor scales mapping absmax to 448 (fp8_act_scales of the INT8 calibration)."""
    QUANT_W = staticmethod(lambda w: quantize_weights_e4m3(w))
    FORWARD = "orc_resnet18_fp8_forward"
    CK_DTYPE = np.uint8
