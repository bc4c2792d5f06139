"""Deterministic synthetic ResNet-18 weights / inputs in the reference's export naming
(<state_dict key>.bin, tools/export_resnet18.py:85-92; names consumed at runtime/infer_e2e.cu:262-330,428-429).

Everything is on power-of-two lattices so C, CUDA and NumPy agree bit-for-bit:
  conv / fc weights   q * 2^-e,  q uniform int in [-127,127], e chosen per layer so the activations keep
                      roughly unit scale through the network (He-style: std(w) ~ sqrt(2/fan_in))
  BN gamma, var       [128,384] * 2^-8  (0.5 .. 1.5);  beta, mean  [-64,64] * 2^-8
  input image         k/64, k in [-192,192]   (covers the reference input's observed range)
The activation scales for these weights are calibrated offline with the FP32 oracle
(tests/golden/make_golden.py) and shipped as synth_act_scales.json.
"""
from __future__ import annotations

import json
import math
import os
from typing import Dict, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

BLOCKS = [(64, 64, 1, False), (64, 64, 1, False), (64, 128, 2, True), (128, 128, 1, False),
          (128, 256, 2, True), (256, 256, 1, False), (256, 512, 2, True), (512, 512, 1, False)]


def conv_keys() -> Dict[int, Tuple[str, str]]:
    """conv index (dlq.h DLQ_NUM_CONVS ordering) -> (weight key, bn key prefix)"""
    keys = {0: ("conv1.weight", "bn1")}
    for b in range(8):
        p = f"layer{b // 2 + 1}.{b % 2}"
        keys[1 + 3 * b] = (p + ".conv1.weight", p + ".bn1")
        keys[2 + 3 * b] = (p + ".conv2.weight", p + ".bn2")
        keys[3 + 3 * b] = (p + ".downsample.0.weight", p + ".downsample.1")
    return keys


def conv_geometry() -> Dict[int, Tuple[int, int, int, int, int]]:
    """conv index -> (ic, oc, k, stride, pad) as wired by runtime/infer_e2e.cu:258-407"""
    g = {0: (3, 64, 7, 2, 3)}
    for b, (ic, oc, s, down) in enumerate(BLOCKS):
        g[1 + 3 * b] = (ic, oc, 3, s, 1)
        g[2 + 3 * b] = (oc, oc, 3, 1, 1)
        if down:
            g[3 + 3 * b] = (ic, oc, 1, s, 0)
    return g


def weight_shift(fan_in: int) -> int:
    return int(round(math.log2(73.3 * math.sqrt(fan_in / 2.0))))


def make_weights(seed: int, fill=None) -> Dict[str, np.ndarray]:
    """fill(shape, seed, name, lo, hi, shift) -> np.float32 array; defaults to the library's generator."""
    if fill is None:
        from . import synth_fill_f32 as fill
    w: Dict[str, np.ndarray] = {}
    keys = conv_keys()
    for idx, (ic, oc, k, s, p) in conv_geometry().items():
        wkey, bn = keys[idx]
        w[wkey] = fill((oc, ic, k, k), seed, wkey, -127, 127, weight_shift(ic * k * k))
        w[bn + ".weight"] = fill((oc,), seed, bn + ".weight", 128, 384, 8)
        w[bn + ".bias"] = fill((oc,), seed, bn + ".bias", -64, 64, 8)
        w[bn + ".running_mean"] = fill((oc,), seed, bn + ".running_mean", -64, 64, 8)
        w[bn + ".running_var"] = fill((oc,), seed, bn + ".running_var", 128, 384, 8)
    w["fc.weight"] = fill((1000, 512), seed, "fc.weight", -127, 127, weight_shift(512))
    w["fc.bias"] = fill((1000,), seed, "fc.bias", -64, 64, 8)
    return w


def make_input(seed: int, n: int, fill=None) -> np.ndarray:
    if fill is None:
        from . import synth_fill_f32 as fill
    return fill((n, 3, 224, 224), seed, "input", -192, 192, 6)


def load_act_scales(seed: int = 0) -> np.ndarray:
    """Activation scales calibrated for make_weights(seed) (absmax/127 of the FP32 oracle on make_input(seed, 8))."""
    with open(os.path.join(_HERE, "synth_act_scales.json")) as f:
        table = json.load(f)
    return np.asarray(table[str(seed)], dtype=np.float32)
