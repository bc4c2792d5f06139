"""Regenerates the committed golden fixtures.  Run in the BUILD container (needs /root/reference):

    python tests/golden/make_golden.py

1. ref_gapfc_*.bin   — the only numeric known-answer set in the reference tree:
                       RKL/tmp_e2e/{l4,gap,fc.weight,fc.bias}.bin -> RKL/out/step8_logits.bin.
                       Data files copied verbatim (fc.weight stored as float16-exact? no: full fp32).
2. dlq_b200/synth_act_scales.json — activation scales for the synthetic weights (seeds 0,1,2), calibrated
                       with the FP32 oracle on make_input(seed=0, n=8): absmax/127 per tensor.
3. i8_seed0_n2.npz   — INT8 oracle outputs (logits + checkpoint digests) for seed 0, 2 images: guards the
                       oracle itself against drift and lets the GPU test compare against a committed vector.
4. mnist_v3_seed.npz — output of the reference's own MN/v3.c forward_timed (oracle/_ref/libref_mnist_v3.so)
                       on a small synthetic batch: pins orc_mnist_mlp_forward.
"""
import ctypes as C
import hashlib
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc  # noqa: E402
from dlq_b200 import synth  # noqa: E402

RKL = "/root/reference/CUDA/resnet18-kernel-lab"


def copy_reference_kat():
    for src, dst in [("tmp_e2e/l4.bin", "ref_gapfc_l4.bin"), ("tmp_e2e/gap.bin", "ref_gapfc_gap.bin"),
                     ("tmp_e2e/fc.weight.bin", "ref_gapfc_fc_weight.bin"), ("tmp_e2e/fc.bias.bin", "ref_gapfc_fc_bias.bin"),
                     ("out/step8_logits.bin", "ref_gapfc_logits.bin")]:
        shutil.copyfile(os.path.join(RKL, src), os.path.join(HERE, dst))


def calibrate():
    table = {}
    x = synth.make_input(0, 8, fill=orc.fill_f32)
    for seed in (0, 1, 2):
        w = synth.make_weights(seed, fill=orc.fill_f32)
        s = orc.calibrate(w, x)
        table[str(seed)] = [float(v) for v in s]
        print("seed", seed, "scales", np.array2string(s, precision=4))
    with open(os.path.join(ROOT, "dlq_b200", "synth_act_scales.json"), "w") as f:
        json.dump(table, f, indent=1)
    return table


def i8_golden(table):
    w = synth.make_weights(0, fill=orc.fill_f32)
    x = synth.make_input(0, 2, fill=orc.fill_f32)
    m = orc.I8Model(w, np.asarray(table["0"], dtype=np.float32))
    out = m.forward(x, checkpoints=True)
    dig = {k: hashlib.sha256(v.tobytes()).hexdigest() for k, v in out.items() if k != "logits"}
    np.savez(os.path.join(HERE, "i8_seed0_n2.npz"), logits=out["logits"], gap=out["gap"],
             digests=json.dumps(dig))
    print("int8 golden logits[0,:5] =", out["logits"][0, :5], "argmax", out["logits"].argmax(1))


def mnist_golden():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_mnist_v3.so"))

    class NN(C.Structure):   # MN/v3.c:41-50 NeuralNetwork
        _fields_ = [(n, C.POINTER(C.c_float)) for n in
                    ("weights1", "weights2", "bias1", "bias2", "grad_weights1", "grad_weights2", "grad_bias1", "grad_bias2")]

    batch, ind, hid, outd = 16, 784, 256, 10
    rng = np.random.default_rng(12345)
    x = ((rng.random((batch, ind), dtype=np.float32) - np.float32(0.1307)) / np.float32(0.3081)).astype(np.float32)
    w1 = (rng.standard_normal((ind, hid)) * np.sqrt(2.0 / ind)).astype(np.float32)
    w2 = (rng.standard_normal((hid, outd)) * np.sqrt(2.0 / hid)).astype(np.float32)
    b1 = (rng.standard_normal(hid) * 0.1).astype(np.float32)
    b2 = (rng.standard_normal(outd) * 0.1).astype(np.float32)
    nn = NN()
    fp = C.POINTER(C.c_float)
    nn.weights1, nn.weights2 = w1.ctypes.data_as(fp), w2.ctypes.data_as(fp)
    nn.bias1, nn.bias2 = b1.ctypes.data_as(fp), b2.ctypes.data_as(fp)
    hidden = np.zeros((batch, hid), dtype=np.float32)
    out = np.zeros((batch, outd), dtype=np.float32)
    stats = (C.c_double * 32)()
    lib.forward_timed(C.byref(nn), x.ctypes.data_as(fp), hidden.ctypes.data_as(fp), out.ctypes.data_as(fp), batch, stats)
    np.savez(os.path.join(HERE, "mnist_v3_seed.npz"), x=x, w1=w1, b1=b1, w2=w2, b2=b2, hidden=hidden, out=out)
    print("mnist v3 reference out[0] =", out[0])


if __name__ == "__main__":
    copy_reference_kat()
    mnist_golden()
    t = calibrate()
    i8_golden(t)
