"""MNIST MLP forward (784 -> 256 -> 10) on the B200 through the C-ABI operators — the GPU counterpart of the
reference's CPU forward (CUDA/MNIST_on_GPU/v3.c:177-215 forward_timed: X@W1 + b1 -> ReLU -> @W2 + b2 -> softmax),
SURVEY 8f-4.  INT8 weights per output row / per-tensor activations (QUANT_SPEC 1-2, 5):

    x_q  = quantise(x, s_x)                                   dlq_quantize_f32_i8
    h    = fmaf(float(x_q . W1q[o]), s_x*s_w1[o], b1[o])      dlq_fc_forward_i8        (v3.c: matmul_a_b + bias_forward)
    h    = relu(h)                                            dlq_relu_forward_f32     (v3.c:161-165)
    h_q  = quantise(h, s_h)                                   dlq_quantize_f32_i8
    z    = fmaf(float(h_q . W2q[o]), s_h*s_w2[o], b2[o])      dlq_fc_forward_i8
    p    = softmax(z)                                         dlq_softmax_f32          (v3.c:108-123, without its 1e-7 clamp)

The reference stores weights as [in, out] (v3.c matmul_a_b(A[m,n], B[n,k])); they are transposed to [out, in] rows
at construction.  Scales: s_x, s_h = absmax / 127 of a calibration batch (numpy, at construction)."""
import numpy as np


def _quant_rows(w):
    """per-row symmetric int8 (QUANT_SPEC 1): s = absmax/127, q = clamp(rne(w * fp32(1/s)), -127, 127)"""
    w = np.ascontiguousarray(w, dtype=np.float32)
    am = np.abs(w).max(axis=1)
    s = np.where(am > 0, (am.astype(np.float64) / 127.0), 1.0).astype(np.float32)
    inv = (1.0 / s.astype(np.float64)).astype(np.float32)
    q = np.clip(np.rint(w * inv[:, None]), -127, 127).astype(np.int8)
    return q, s


class MnistMLP:
    def __init__(self, ctx, w1, b1, w2, b2, x_calib):
        import torch
        self.ctx = ctx
        w1, w2 = np.asarray(w1, np.float32), np.asarray(w2, np.float32)      # [784,256], [256,10] as in v3.c
        self.w1q, self.s_w1 = _quant_rows(w1.T)
        self.w2q, self.s_w2 = _quant_rows(w2.T)
        xc = np.asarray(x_calib, np.float32)
        hc = np.maximum(xc @ w1 + np.asarray(b1, np.float32), 0)
        self.s_x = np.float32(np.float64(np.abs(xc).max()) / 127.0)
        self.s_h = np.float32(np.float64(max(np.abs(hc).max(), 1e-6)) / 127.0)
        self.sc1 = (np.float64(self.s_x) * self.s_w1.astype(np.float64)).astype(np.float32)
        self.sc2 = (np.float64(self.s_h) * self.s_w2.astype(np.float64)).astype(np.float32)
        self.b1, self.b2 = np.asarray(b1, np.float32), np.asarray(b2, np.float32)
        dev = f"cuda:{ctx.device}"
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self.d = {k: t(v) for k, v in dict(w1q=self.w1q, w2q=self.w2q, sc1=self.sc1, sc2=self.sc2, b1=self.b1, b2=self.b2).items()}

    def forward(self, x, probs=None):
        """x: float32 [B,784] device tensor -> (logits [B,10], probs [B,10]) device tensors"""
        import torch
        B = x.shape[0]
        dev = x.device
        xq = torch.empty((B, 784), dtype=torch.int8, device=dev)
        h = torch.empty((B, 256), dtype=torch.float32, device=dev)
        hq = torch.empty((B, 256), dtype=torch.int8, device=dev)
        z = torch.empty((B, 10), dtype=torch.float32, device=dev)
        p = probs if probs is not None else torch.empty((B, 10), dtype=torch.float32, device=dev)
        c, d = self.ctx, self.d
        c.quantize_f32_i8(x, float(self.s_x), xq)
        c.fc_forward_i8(xq, d["w1q"], d["sc1"], d["b1"], h)
        c.relu_forward_f32(h)
        c.quantize_f32_i8(h, float(self.s_h), hq)
        c.fc_forward_i8(hq, d["w2q"], d["sc2"], d["b2"], z)
        c.softmax_f32(z, p)
        return z, p
