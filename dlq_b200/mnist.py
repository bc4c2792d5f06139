"""MNIST MLP forward (784 -> 256 -> 10) on the B200 behind the C ABI (dlq_mlp_*, dlq_b200/csrc/fc_mlp.cu) - the GPU
counterpart of the reference's forward (CUDA/MNIST_on_GPU/v4.cu:255-302 forward_timed, v5.cu:127-157, and v3.c:177-215 on
the CPU: X@W1 + b1 -> ReLU -> @W2 + b2 -> softmax), SURVEY 8f-4.  Both FC layers run on the tcgen05 GEMM core (the conv
kernel as a 1x1 convolution); arithmetic in include/dlq.h (dlq_mlp_create).

This module only calibrates the two activation scales (absmax / 127, or / 448 for E4M3, of a calibration batch - the
PTQ recipe of QUANT_SPEC 5) and forwards to the C entry points."""
import numpy as np

from . import MLP


def calibrate(w1, b1, w2, b2, x_calib, fp8: bool = False):
    """(s_x, s_h): absmax / qmax of the input and of the post-ReLU hidden activations on x_calib, in float32"""
    xc = np.asarray(x_calib, np.float32)
    hc = np.maximum(xc @ np.asarray(w1, np.float32) + np.asarray(b1, np.float32), 0)
    qmax = 448.0 if fp8 else 127.0
    s_x = np.float32(np.float64(np.abs(xc).max()) / qmax)
    s_h = np.float32(np.float64(max(np.abs(hc).max(), 1e-6)) / qmax)
    return s_x, s_h


class MnistMLP(MLP):
    def __init__(self, ctx, w1, b1, w2, b2, x_calib, max_batch: int = 1024, fp8: bool = False):
        self.s_x, self.s_h = calibrate(w1, b1, w2, b2, x_calib, fp8)
        super().__init__(ctx, w1, b1, w2, b2, self.s_x, self.s_h, max_batch, fp8)
