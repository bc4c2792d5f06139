"""dlq_b200 — host-side Python mirror of the libdlq_b200.so C ABI (include/dlq.h).

The product is the CUDA library; this module only binds it with ctypes so that tests and bench.py can
drive it with torch tensors as device memory.  There is NO CPU fallback: importing works without a
GPU (so the ABI can be inspected), but creating a Context without the library or without a B200
raises.

Reference surface mirrored here (yeontachi/DLQ, CUDA/resnet18-kernel-lab/cpp/fp32):
  conv2d_nchw_im2col_gemm   runtime/infer_e2e.cu:102-136   -> Context.conv2d_i8
  bn_launch / bn_inference  runtime/infer_e2e.cu:83-97     -> Context.bn_inference_f32 (standalone) / folded
  relu_forward, add_inplace kernels/relu.cu, kernels/add.cu
  maxpool2d_3x3_s2p1_nchw   kernels/maxpool2d.cu:5-41
  gap_global, fc_forward    kernels/gap_global.cu, runtime/infer_e2e.cu:206-219
  main() wiring             runtime/infer_e2e.cu:254-433   -> ResNet18.forward
"""
from __future__ import annotations

import ctypes as C
import functools
import os
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# (DLQ_B200_LIB: tuning tools load the `make TIMING=1` build - same ABI, per-role cycle counters and experiment switches)
LIB_PATH = os.environ.get("DLQ_B200_LIB") or os.path.join(_HERE, "libdlq_b200.so")

NUM_CONVS = 25
NUM_ACTS = 27
ACT_INPUT, ACT_STEM, ACT_GAP = 0, 1, 26

# every symbol include/dlq.h declares (checked by tests/test_abi.py against the header and the .so)
ABI_SYMBOLS = [
    "dlq_create", "dlq_destroy", "dlq_last_error_string", "dlq_sync", "dlq_stream", "dlq_set_stream", "dlq_version",
    "dlq_quantize_f32_i8", "dlq_dequantize_i8_f32", "dlq_dequantize_i8_f32_per_channel",
    "dlq_quantize_f32_e4m3", "dlq_dequantize_e4m3_f32", "dlq_conv_weights_pack_fp8", "dlq_conv_weights_pack_e4m3",
    "dlq_f32_to_e4m3", "dlq_conv2d_fp8",
    "dlq_conv_weights_pack", "dlq_conv_weights_pack_i8", "dlq_conv_weights_free", "dlq_conv2d_i8", "dlq_fold_bn",
    "dlq_res_mul", "dlq_act_bytes", "dlq_conv_required_pad_rows", "dlq_conv2d_i8_act", "dlq_act_from_nchw_i8",
    "dlq_act_to_nchw_i8", "dlq_stem_pack_input_i8", "dlq_conv_plan_create", "dlq_conv_plan_launch",
    "dlq_conv_plan_destroy",
    "dlq_bn_inference_f32", "dlq_relu_forward_f32", "dlq_relu_forward_i8", "dlq_add_inplace_f32", "dlq_add_requant_i8",
    "dlq_maxpool2d_3x3_s2p1_nchw_i8", "dlq_gap_global_i8", "dlq_fc_forward_i8", "dlq_softmax_f32",
    "dlq_resnet18_create", "dlq_resnet18_destroy", "dlq_resnet18_forward", "dlq_resnet18_forward_host",
    "dlq_resnet18_set_preprocess", "dlq_resnet18_forward_u8", "dlq_resnet18_forward_host_u8",
    "dlq_resnet18_checkpoint", "dlq_resnet18_graph_capture", "dlq_resnet18_graph_launch", "dlq_resnet18_launches", "dlq_resnet18_profile", "dlq_synth_fill_f32",
    "dlq_multi_create", "dlq_multi_destroy", "dlq_multi_forward_host", "dlq_multi_last_error_string",
    "dlq_weight_dir_load", "dlq_weight_dir_weights", "dlq_weight_dir_free", "dlq_weight_dir_save",
    "dlq_resnet18_f32_create", "dlq_resnet18_f32_destroy", "dlq_resnet18_f32_forward", "dlq_resnet18_f32_checkpoint",
    "dlq_resnet18_f32_absmax", "dlq_resnet18_f32_reset_absmax", "dlq_act_scales_from_absmax",
    "dlq_topk_f32", "dlq_compare_f32",
    "dlq_workspace_reserve", "dlq_workspace_bytes", "dlq_conv2d_workspace_bytes",
    "dlq_fc_weights_pack", "dlq_fc_weights_pack_i8", "dlq_fc_weights_pack_e4m3", "dlq_fc_workspace_bytes",
    "dlq_fc_forward_i8_tc", "dlq_fc_forward_fp8",
    "dlq_resnet18_submit_host", "dlq_resnet18_submit_host_u8", "dlq_resnet18_wait", "dlq_resnet18_launches_for_batch",
    "dlq_resnet18_enable_stamps", "dlq_resnet18_read_stamps", "dlq_resnet18_set_option", "dlq_resnet18_dep_timeouts", "dlq_resnet18_plan_info",
    "dlq_multi_n_devices", "dlq_multi_set_preprocess", "dlq_multi_forward_host_u8", "dlq_multi_submit_host",
    "dlq_multi_submit_host_u8", "dlq_multi_wait", "dlq_multi_forward_device",
    "dlq_mlp_create", "dlq_mlp_destroy", "dlq_mlp_forward", "dlq_mlp_checkpoint", "dlq_mlp_weight_scales",
]


class DlqError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"dlq error {code}: {msg}")
        self.code = code


class _Epilogue(C.Structure):
    _fields_ = [("alpha", C.c_void_p), ("beta", C.c_void_p), ("residual", C.c_void_p), ("res_mul", C.c_float),
                ("relu", C.c_int)]


class _Act(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("N", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int), ("PR", C.c_int)]


class _ResNet18Weights(C.Structure):
    _fields_ = [("conv_w", C.c_void_p * NUM_CONVS), ("bn_gamma", C.c_void_p * NUM_CONVS),
                ("bn_beta", C.c_void_p * NUM_CONVS), ("bn_mean", C.c_void_p * NUM_CONVS),
                ("bn_var", C.c_void_p * NUM_CONVS), ("fc_w", C.c_void_p), ("fc_b", C.c_void_p),
                ("act_scale", C.c_float * NUM_ACTS), ("fp8", C.c_int)]


_lib = None


def load_library() -> C.CDLL:
    """Load libdlq_b200.so; raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} not built; run `make -C dlq_b200/csrc` (no CPU fallback exists)")
    lib = C.CDLL(LIB_PATH)
    vp, i, f, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    sig = {
        "dlq_create": (i, [i, C.POINTER(vp)]),
        "dlq_destroy": (None, [vp]),
        "dlq_last_error_string": (C.c_char_p, [vp]),
        "dlq_sync": (i, [vp]),
        "dlq_stream": (vp, [vp]),
        "dlq_set_stream": (i, [vp, vp]),
        "dlq_version": (C.c_char_p, []),
        "dlq_quantize_f32_i8": (i, [vp, vp, sz, f, vp]),
        "dlq_dequantize_i8_f32": (i, [vp, vp, sz, f, vp]),
        "dlq_dequantize_i8_f32_per_channel": (i, [vp, vp, i, i, i, vp, vp]),
        "dlq_quantize_f32_e4m3": (i, [vp, vp, sz, f, vp]),
        "dlq_dequantize_e4m3_f32": (i, [vp, vp, sz, f, vp]),
        "dlq_conv_weights_pack_fp8": (i, [vp, vp, i, i, i, i, i, i, i, i, vp, C.POINTER(vp)]),
        "dlq_conv_weights_pack_e4m3": (i, [vp, vp, i, i, i, i, i, i, i, i, C.POINTER(vp)]),
        "dlq_f32_to_e4m3": (C.c_uint8, [f]),
        "dlq_conv2d_fp8": (i, [vp, vp, i, i, i, i, vp, C.POINTER(_Epilogue), vp, vp, C.POINTER(i), C.POINTER(i)]),
        "dlq_conv_weights_pack": (i, [vp, vp, i, i, i, i, i, i, i, i, vp, C.POINTER(vp)]),
        "dlq_conv_weights_pack_i8": (i, [vp, vp, i, i, i, i, i, i, i, i, C.POINTER(vp)]),
        "dlq_conv_weights_free": (None, [vp]),
        "dlq_conv2d_i8": (i, [vp, vp, i, i, i, i, vp, C.POINTER(_Epilogue), vp, vp, C.POINTER(i), C.POINTER(i)]),
        "dlq_fold_bn": (None, [vp, vp, vp, vp, f, vp, f, f, i, vp, vp]),
        "dlq_res_mul": (f, [f, f]),
        "dlq_act_bytes": (sz, [i, i, i, i, i]),
        "dlq_conv_required_pad_rows": (i, [vp]),
        "dlq_conv2d_i8_act": (i, [vp, C.POINTER(_Act), vp, C.POINTER(_Epilogue), C.POINTER(_Act), C.POINTER(_Act), vp]),
        "dlq_conv_plan_create": (i, [vp, C.POINTER(_Act), vp, C.POINTER(_Epilogue), C.POINTER(_Act), C.POINTER(_Act), vp,
                                     C.POINTER(vp)]),
        "dlq_conv_plan_launch": (i, [vp, vp]),
        "dlq_conv_plan_destroy": (None, [vp]),
        "dlq_act_from_nchw_i8": (i, [vp, vp, C.POINTER(_Act)]),
        "dlq_act_to_nchw_i8": (i, [vp, C.POINTER(_Act), vp]),
        "dlq_stem_pack_input_i8": (i, [vp, vp, i, i, i, C.POINTER(_Act)]),
        "dlq_bn_inference_f32": (i, [vp, vp, vp, vp, vp, vp, f, i, i, i, i]),
        "dlq_relu_forward_f32": (i, [vp, vp, sz]),
        "dlq_relu_forward_i8": (i, [vp, vp, sz]),
        "dlq_add_inplace_f32": (i, [vp, vp, vp, sz]),
        "dlq_add_requant_i8": (i, [vp, vp, f, vp, f, sz, i, f]),
        "dlq_maxpool2d_3x3_s2p1_nchw_i8": (i, [vp, vp, i, i, i, i, vp]),
        "dlq_gap_global_i8": (i, [vp, vp, i, i, i, i, f, f, vp, vp]),
        "dlq_fc_forward_i8": (i, [vp, vp, vp, vp, vp, i, i, i, vp]),
        "dlq_softmax_f32": (i, [vp, vp, i, i, vp]),
        "dlq_resnet18_create": (i, [vp, C.POINTER(_ResNet18Weights), i, C.POINTER(vp)]),
        "dlq_resnet18_destroy": (None, [vp]),
        "dlq_resnet18_forward": (i, [vp, vp, i, vp]),
        "dlq_resnet18_forward_host": (i, [vp, vp, i, vp]),
        "dlq_resnet18_checkpoint": (i, [vp, C.c_char_p, vp]),
        "dlq_resnet18_graph_capture": (i, [vp, vp, i, vp]),
        "dlq_resnet18_set_preprocess": (i, [vp, vp, vp]),
        "dlq_resnet18_forward_u8": (i, [vp, vp, i, vp]),
        "dlq_resnet18_forward_host_u8": (i, [vp, vp, i, vp]),
        "dlq_resnet18_graph_launch": (i, [vp]),
        "dlq_resnet18_launches": (i, [vp]),
        "dlq_resnet18_profile": (i, [vp, vp, i, vp, vp]),
        "dlq_synth_fill_f32": (None, [vp, sz, C.c_uint64, C.c_char_p, i, i, i]),
        "dlq_multi_create": (i, [C.POINTER(i), i, C.POINTER(_ResNet18Weights), i, C.POINTER(vp)]),
        "dlq_multi_destroy": (None, [vp]),
        "dlq_multi_forward_host": (i, [vp, vp, i, vp]),
        "dlq_multi_last_error_string": (C.c_char_p, [vp]),
        "dlq_weight_dir_load": (i, [C.c_char_p, C.POINTER(vp), C.c_char_p, sz]),
        "dlq_weight_dir_weights": (C.POINTER(_ResNet18Weights), [vp]),
        "dlq_weight_dir_free": (None, [vp]),
        "dlq_weight_dir_save": (i, [C.c_char_p, C.POINTER(_ResNet18Weights), i]),
        "dlq_resnet18_f32_create": (i, [vp, C.POINTER(_ResNet18Weights), i, C.POINTER(vp)]),
        "dlq_resnet18_f32_destroy": (None, [vp]),
        "dlq_resnet18_f32_forward": (i, [vp, vp, i, vp]),
        "dlq_resnet18_f32_checkpoint": (i, [vp, C.c_char_p, vp]),
        "dlq_resnet18_f32_absmax": (i, [vp, vp]),
        "dlq_resnet18_f32_reset_absmax": (i, [vp]),
        "dlq_act_scales_from_absmax": (None, [vp, i, vp]),
        "dlq_topk_f32": (i, [vp, vp, i, i, i, vp, vp]),
        "dlq_compare_f32": (i, [vp, vp, vp, sz, vp]),
        "dlq_workspace_reserve": (i, [vp, sz]),
        "dlq_workspace_bytes": (sz, [vp]),
        "dlq_conv2d_workspace_bytes": (sz, [vp, i, i, i, i, i]),
        "dlq_fc_weights_pack": (i, [vp, vp, i, i, i, vp, C.POINTER(vp)]),
        "dlq_fc_weights_pack_i8": (i, [vp, vp, i, i, C.POINTER(vp)]),
        "dlq_fc_weights_pack_e4m3": (i, [vp, vp, i, i, C.POINTER(vp)]),
        "dlq_fc_workspace_bytes": (sz, [vp, i]),
        "dlq_fc_forward_i8_tc": (i, [vp, vp, vp, vp, vp, i, vp]),
        "dlq_fc_forward_fp8": (i, [vp, vp, vp, vp, vp, i, vp]),
        "dlq_resnet18_submit_host": (i, [vp, vp, i, vp]),
        "dlq_resnet18_submit_host_u8": (i, [vp, vp, i, vp]),
        "dlq_resnet18_wait": (i, [vp]),
        "dlq_resnet18_launches_for_batch": (i, [vp, i]),
        "dlq_resnet18_enable_stamps": (i, [vp, i]),
        "dlq_resnet18_read_stamps": (i, [vp, vp, C.POINTER(i)]),
        "dlq_resnet18_set_option": (i, [vp, C.c_char_p, i]),
        "dlq_resnet18_dep_timeouts": (i, [vp, C.POINTER(C.c_uint)]),
        "dlq_resnet18_plan_info": (i, [vp, i, C.c_char_p, C.POINTER(i)]),
        "dlq_multi_n_devices": (i, [vp]),
        "dlq_multi_set_preprocess": (i, [vp, vp, vp]),
        "dlq_multi_forward_host_u8": (i, [vp, vp, i, vp]),
        "dlq_multi_submit_host": (i, [vp, vp, i, vp]),
        "dlq_multi_submit_host_u8": (i, [vp, vp, i, vp]),
        "dlq_multi_wait": (i, [vp]),
        "dlq_multi_forward_device": (i, [vp, C.POINTER(vp), C.POINTER(i), C.POINTER(vp)]),
        "dlq_mlp_create": (i, [vp, vp, vp, vp, vp, i, i, i, f, f, i, i, C.POINTER(vp)]),
        "dlq_mlp_destroy": (None, [vp]),
        "dlq_mlp_forward": (i, [vp, vp, i, vp, vp]),
        "dlq_mlp_checkpoint": (i, [vp, C.c_char_p, vp]),
        "dlq_mlp_weight_scales": (i, [vp, i, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _ptr(t) -> Optional[int]:
    """device pointer of a torch tensor (or None)"""
    if t is None:
        return None
    return t.data_ptr()


def synth_fill_f32(shape, seed: int, name: str, lo: int, hi: int, shift: int) -> np.ndarray:
    """Deterministic lattice tensor: (lo + splitmix64(seed,name) % (hi-lo+1)) * 2^-shift."""
    lib = load_library()
    a = np.empty(shape, dtype=np.float32)
    lib.dlq_synth_fill_f32(a.ctypes.data, a.size, seed, name.encode(), lo, hi, shift)
    return a


def fold_bn(g, b, m, v, s_w, s_x: float, s_y: float, eps: float = 1e-5):
    """host: (alpha, beta) requantisation constants per QUANT_SPEC §3"""
    lib = load_library()
    arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (g, b, m, v, s_w)]
    oc = arrs[0].size
    alpha, beta = np.empty(oc, np.float32), np.empty(oc, np.float32)
    lib.dlq_fold_bn(arrs[0].ctypes.data, arrs[1].ctypes.data, arrs[2].ctypes.data, arrs[3].ctypes.data, eps,
                    arrs[4].ctypes.data, s_x, s_y, oc, alpha.ctypes.data, beta.ctypes.data)
    return alpha, beta


def res_mul(s_r: float, s_y: float) -> float:
    return float(load_library().dlq_res_mul(s_r, s_y))


class ConvWeights:
    def __init__(self, ctx: "Context", handle: int, scale: Optional[np.ndarray], oc: int):
        self.ctx, self.handle, self.scale, self.oc = ctx, handle, scale, oc

    def free(self):
        if self.handle:
            load_library().dlq_conv_weights_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# Context methods that launch kernels reading / writing caller tensors (ordered after torch's stream, see Context)
_ORDERED_METHODS = frozenset([
    "quantize_f32_i8", "dequantize_i8_f32", "dequantize_i8_f32_per_channel", "quantize_f32_e4m3", "dequantize_e4m3_f32",
    "conv2d_fp8", "conv2d_i8", "conv2d_i8_act", "conv_plan", "act_from_nchw_i8", "act_to_nchw_i8", "stem_pack_input_i8",
    "bn_inference_f32", "relu_forward_f32", "relu_forward_i8", "add_inplace_f32", "add_requant_i8",
    "maxpool2d_3x3_s2p1_nchw_i8", "gap_global_i8", "fc_forward_i8", "fc_forward_i8_tc", "fc_forward_fp8", "softmax_f32",
    "topk_f32", "compare_f32",
])


class Context:
    """One per device; enqueues on its own NON-BLOCKING stream (see dlq.h), which has no implicit ordering with torch's
    streams.  Tensors made by torch (`torch.zeros`, `.cuda()`, ...) are produced on torch's current stream, so every
    call of this mirror first makes the library's stream wait for torch's current stream (`order_after_torch`, an event
    record + stream wait; a few microseconds of host time).  A timing loop whose buffers are already complete switches
    that off with `auto_order = False`."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.dlq_create(device, C.byref(h))
        if rc != 0:
            raise DlqError(rc, f"dlq_create(device={device}) failed: no usable sm_100 GPU (there is no CPU fallback)")
        self.h = h
        self.device = device
        self.auto_order = True
        self._ext = None

    def order_after_torch(self):
        """work enqueued by the library from here on runs after everything torch has enqueued on its current stream"""
        if not self.auto_order:
            return
        import torch
        if self._ext is None:
            self._ext = torch.cuda.ExternalStream(self.stream, device=self.device)
        self._ext.wait_stream(torch.cuda.current_stream(self.device))

    def close(self):
        if self.h:
            self.lib.dlq_destroy(self.h)
            self.h = None

    def _ck(self, rc: int):
        if rc != 0:
            raise DlqError(rc, self.lib.dlq_last_error_string(self.h).decode())

    def sync(self):
        self._ck(self.lib.dlq_sync(self.h))

    @property
    def stream(self) -> int:
        return self.lib.dlq_stream(self.h)

    # ---- quantise / dequantise
    def quantize_f32_i8(self, x, scale: float, q):
        self._ck(self.lib.dlq_quantize_f32_i8(self.h, _ptr(x), x.numel(), scale, _ptr(q)))

    def dequantize_i8_f32(self, q, scale: float, x):
        self._ck(self.lib.dlq_dequantize_i8_f32(self.h, _ptr(q), q.numel(), scale, _ptr(x)))

    def dequantize_i8_f32_per_channel(self, q, scales, x):
        n, c = q.shape[0], q.shape[1]
        hw = q.numel() // max(1, n * c)
        self._ck(self.lib.dlq_dequantize_i8_f32_per_channel(self.h, _ptr(q), n, c, hw, _ptr(scales), _ptr(x)))

    # ---- conv
    def pack_conv_weights(self, w_oihw: np.ndarray, stride: int, pad: int) -> ConvWeights:
        w = np.ascontiguousarray(w_oihw, dtype=np.float32)
        oc, ic, kh, kw = w.shape
        scale = np.empty(oc, dtype=np.float32)
        h = C.c_void_p()
        self._ck(self.lib.dlq_conv_weights_pack(self.h, w.ctypes.data, oc, ic, kh, kw, stride, stride, pad, pad,
                                                scale.ctypes.data, C.byref(h)))
        return ConvWeights(self, h, scale, oc)

    def pack_conv_weights_i8(self, wq_oihw: np.ndarray, stride: int, pad: int) -> ConvWeights:
        w = np.ascontiguousarray(wq_oihw, dtype=np.int8)
        oc, ic, kh, kw = w.shape
        h = C.c_void_p()
        self._ck(self.lib.dlq_conv_weights_pack_i8(self.h, w.ctypes.data, oc, ic, kh, kw, stride, stride, pad, pad,
                                                   C.byref(h)))
        return ConvWeights(self, h, None, oc)

    # ---- FP8 (E4M3) variants, QUANT_SPEC section 6
    def quantize_f32_e4m3(self, x, scale: float, q):
        self._ck(self.lib.dlq_quantize_f32_e4m3(self.h, _ptr(x), x.numel(), scale, _ptr(q)))

    def dequantize_e4m3_f32(self, q, scale: float, x):
        self._ck(self.lib.dlq_dequantize_e4m3_f32(self.h, _ptr(q), q.numel(), scale, _ptr(x)))

    def pack_conv_weights_fp8(self, w_oihw: np.ndarray, stride: int, pad: int) -> ConvWeights:
        w = np.ascontiguousarray(w_oihw, dtype=np.float32)
        oc, ic, kh, kw = w.shape
        scale = np.empty(oc, dtype=np.float32)
        h = C.c_void_p()
        self._ck(self.lib.dlq_conv_weights_pack_fp8(self.h, w.ctypes.data, oc, ic, kh, kw, stride, stride, pad, pad,
                                                    scale.ctypes.data, C.byref(h)))
        return ConvWeights(self, h, scale, oc)

    def pack_conv_weights_e4m3(self, wq_oihw: np.ndarray, stride: int, pad: int) -> ConvWeights:
        w = np.ascontiguousarray(wq_oihw, dtype=np.uint8)
        oc, ic, kh, kw = w.shape
        h = C.c_void_p()
        self._ck(self.lib.dlq_conv_weights_pack_e4m3(self.h, w.ctypes.data, oc, ic, kh, kw, stride, stride, pad, pad,
                                                     C.byref(h)))
        return ConvWeights(self, h, None, oc)

    def conv2d_fp8(self, x, w: ConvWeights, alpha=None, beta=None, residual=None, res_mul: float = 0.0,
                   relu: bool = False, y=None, acc_out=None):
        """x / y / residual: uint8 E4M3 codes NCHW; acc_out: float32 NCHW raw accumulators"""
        n, c, hh, ww = x.shape
        ep = _Epilogue(_ptr(alpha), _ptr(beta), _ptr(residual), res_mul, int(relu))
        oh, ow = C.c_int(), C.c_int()
        self._ck(self.lib.dlq_conv2d_fp8(self.h, _ptr(x), n, c, hh, ww, w.handle,
                                         C.byref(ep) if alpha is not None else None, _ptr(y), _ptr(acc_out),
                                         C.byref(oh), C.byref(ow)))
        return oh.value, ow.value

    def conv2d_i8(self, x, w: ConvWeights, alpha=None, beta=None, residual=None, res_mul: float = 0.0,
                  relu: bool = False, y=None, acc_out=None):
        """alpha/beta/res_mul are requantisation multipliers (output scale folded in, see fold_bn)."""
        n, c, hh, ww = x.shape
        ep = _Epilogue(_ptr(alpha), _ptr(beta), _ptr(residual), res_mul, int(relu))
        oh, ow = C.c_int(), C.c_int()
        self._ck(self.lib.dlq_conv2d_i8(self.h, _ptr(x), n, c, hh, ww, w.handle,
                                        C.byref(ep) if alpha is not None else None, _ptr(y), _ptr(acc_out),
                                        C.byref(oh), C.byref(ow)))
        return oh.value, ow.value

    # ---- native layout (row-padded NHWC) entry points
    def new_act(self, n, h, w, c, pr):
        """allocate a zero-filled row-padded NHWC int8 tensor; returns (torch buffer, _Act)"""
        import torch
        nbytes = self.lib.dlq_act_bytes(n, h, w, c, pr)
        buf = torch.zeros(nbytes + 1024, dtype=torch.int8, device=f"cuda:{self.device}")
        self.order_after_torch()      # the zero fill runs on torch's stream
        return buf, _Act(buf.data_ptr(), n, h, w, c, pr)

    def required_pad_rows(self, w: ConvWeights) -> int:
        return self.lib.dlq_conv_required_pad_rows(w.handle)

    def conv2d_i8_act(self, x_act, w: ConvWeights, y_act=None, alpha=None, beta=None, residual_act=None,
                      res_mul: float = 0.0, relu: bool = False, acc_out=None):
        ep = _Epilogue(_ptr(alpha), _ptr(beta), None, res_mul, int(relu))
        self._ck(self.lib.dlq_conv2d_i8_act(self.h, C.byref(x_act), w.handle, C.byref(ep) if alpha is not None else None,
                                            C.byref(residual_act) if residual_act is not None else None,
                                            C.byref(y_act) if y_act is not None else None, _ptr(acc_out)))

    def conv_plan(self, x_act, w: ConvWeights, y_act=None, alpha=None, beta=None, residual_act=None,
                  res_mul: float = 0.0, relu: bool = False, acc_out=None):
        """plan once; returns a callable that enqueues the conv (no host planning per launch)"""
        ep = _Epilogue(_ptr(alpha), _ptr(beta), None, res_mul, int(relu))
        h = C.c_void_p()
        self._ck(self.lib.dlq_conv_plan_create(self.h, C.byref(x_act), w.handle, C.byref(ep) if alpha is not None else None,
                                               C.byref(residual_act) if residual_act is not None else None,
                                               C.byref(y_act) if y_act is not None else None, _ptr(acc_out), C.byref(h)))

        def launch():
            self._ck(self.lib.dlq_conv_plan_launch(self.h, h))
        launch.handle = h
        launch.destroy = lambda: self.lib.dlq_conv_plan_destroy(h)
        return launch

    def act_from_nchw_i8(self, x, act):
        self._ck(self.lib.dlq_act_from_nchw_i8(self.h, _ptr(x), C.byref(act)))

    def act_to_nchw_i8(self, act, y):
        self._ck(self.lib.dlq_act_to_nchw_i8(self.h, C.byref(act), _ptr(y)))

    def stem_pack_input_i8(self, x, act):
        n, c, hh, ww = x.shape
        self._ck(self.lib.dlq_stem_pack_input_i8(self.h, _ptr(x), n, hh, ww, C.byref(act)))

    # ---- element-wise / pooling
    def bn_inference_f32(self, x, g, b, m, v, eps: float = 1e-5):
        n, c, oh, ow = x.shape
        self._ck(self.lib.dlq_bn_inference_f32(self.h, _ptr(x), _ptr(g), _ptr(b), _ptr(m), _ptr(v), eps, n, c, oh, ow))

    def relu_forward_f32(self, x):
        self._ck(self.lib.dlq_relu_forward_f32(self.h, _ptr(x), x.numel()))

    def relu_forward_i8(self, x):
        self._ck(self.lib.dlq_relu_forward_i8(self.h, _ptr(x), x.numel()))

    def add_inplace_f32(self, y, x):
        self._ck(self.lib.dlq_add_inplace_f32(self.h, _ptr(y), _ptr(x), y.numel()))

    def add_requant_i8(self, y, y_scale, x, x_scale, relu, out_scale):
        self._ck(self.lib.dlq_add_requant_i8(self.h, _ptr(y), y_scale, _ptr(x), x_scale, y.numel(), int(relu), out_scale))

    def maxpool2d_3x3_s2p1_nchw_i8(self, x, y):
        n, c, hh, ww = x.shape
        self._ck(self.lib.dlq_maxpool2d_3x3_s2p1_nchw_i8(self.h, _ptr(x), n, c, hh, ww, _ptr(y)))

    def gap_global_i8(self, x, in_scale, out_scale, y_f32=None, y_i8=None):
        n, c, hh, ww = x.shape
        self._ck(self.lib.dlq_gap_global_i8(self.h, _ptr(x), n, c, hh, ww, in_scale, out_scale, _ptr(y_f32), _ptr(y_i8)))

    def fc_forward_i8(self, g, w, scale, bias, logits):
        n, i = g.shape
        o = w.shape[0]
        self._ck(self.lib.dlq_fc_forward_i8(self.h, _ptr(g), _ptr(w), _ptr(scale), _ptr(bias), n, o, i, _ptr(logits)))

    # ---- tensor-core FC (the conv core as a 1x1 convolution), SURVEY 8f-4
    def pack_fc_weights(self, w_oi: np.ndarray, fp8: bool = False) -> ConvWeights:
        """w_oi: float32 [O, I] (the reference's fc.weight layout); per-row quantisation, QUANT_SPEC 1 / 6"""
        w = np.ascontiguousarray(w_oi, dtype=np.float32)
        o, i_ = w.shape
        scale = np.empty(o, dtype=np.float32)
        h = C.c_void_p()
        self._ck(self.lib.dlq_fc_weights_pack(self.h, w.ctypes.data, o, i_, 1 if fp8 else 0, scale.ctypes.data, C.byref(h)))
        return ConvWeights(self, h, scale, o)

    def pack_fc_weights_i8(self, wq_oi: np.ndarray) -> ConvWeights:
        w = np.ascontiguousarray(wq_oi, dtype=np.int8)
        h = C.c_void_p()
        self._ck(self.lib.dlq_fc_weights_pack_i8(self.h, w.ctypes.data, w.shape[0], w.shape[1], C.byref(h)))
        return ConvWeights(self, h, None, w.shape[0])

    def pack_fc_weights_e4m3(self, wq_oi: np.ndarray) -> ConvWeights:
        w = np.ascontiguousarray(wq_oi, dtype=np.uint8)
        h = C.c_void_p()
        self._ck(self.lib.dlq_fc_weights_pack_e4m3(self.h, w.ctypes.data, w.shape[0], w.shape[1], C.byref(h)))
        return ConvWeights(self, h, None, w.shape[0])

    def fc_forward_i8_tc(self, g, w: ConvWeights, scale, bias, logits):
        self._ck(self.lib.dlq_fc_forward_i8_tc(self.h, _ptr(g), w.handle, _ptr(scale), _ptr(bias), g.shape[0], _ptr(logits)))

    def fc_forward_fp8(self, g, w: ConvWeights, scale, bias, logits):
        self._ck(self.lib.dlq_fc_forward_fp8(self.h, _ptr(g), w.handle, _ptr(scale), _ptr(bias), g.shape[0], _ptr(logits)))

    def workspace_reserve(self, nbytes: int):
        self._ck(self.lib.dlq_workspace_reserve(self.h, nbytes))

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.dlq_workspace_bytes(self.h))

    def conv2d_workspace_bytes(self, w: ConvWeights, n, h, w_, has_residual=False, want_acc=False) -> int:
        return int(self.lib.dlq_conv2d_workspace_bytes(w.handle, n, h, w_, int(has_residual), int(want_acc)))

    def softmax_f32(self, x, y):
        n, k = x.shape
        self._ck(self.lib.dlq_softmax_f32(self.h, _ptr(x), n, k, _ptr(y)))

    def topk_f32(self, x, k: int, idx, val=None):
        """idx: int32 [N,k] device tensor; val: optional float32 [N,k]"""
        n, kk = x.shape
        self._ck(self.lib.dlq_topk_f32(self.h, _ptr(x), n, kk, k, _ptr(idx), _ptr(val)))

    def compare_f32(self, a, b) -> Dict[str, float]:
        """max_abs / mean_abs / cosine of two equally sized float32 device tensors (R/utils.hpp diff_max_mean +
        tools/diag_e2e_compare.py cosine)"""
        out = np.zeros(3, dtype=np.float64)
        self._ck(self.lib.dlq_compare_f32(self.h, _ptr(a), _ptr(b), a.numel(), out.ctypes.data))
        return {"max_abs": float(out[0]), "mean_abs": float(out[1]), "cosine": float(out[2])}


def _ordered(fn):
    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        self.order_after_torch()
        return fn(self, *args, **kwargs)
    return wrapper


for _name in _ORDERED_METHODS:
    setattr(Context, _name, _ordered(getattr(Context, _name)))


def _weights_struct(weights: Dict[str, np.ndarray], act_scale, fp8: bool = False) -> tuple:
    """Build the dlq_resnet18_weights struct from a dict keyed like the reference's <key>.bin export
    (tools/export_resnet18.py:85-92); returns (struct, keepalive list)."""
    from .synth import conv_keys
    s = _ResNet18Weights()
    keep = []

    def put(arr_field, idx, key):
        a = np.ascontiguousarray(weights[key], dtype=np.float32)
        keep.append(a)
        arr_field[idx] = a.ctypes.data

    for idx, (wkey, bnkey) in conv_keys().items():
        if wkey not in weights:
            continue
        put(s.conv_w, idx, wkey)
        put(s.bn_gamma, idx, bnkey + ".weight")
        put(s.bn_beta, idx, bnkey + ".bias")
        put(s.bn_mean, idx, bnkey + ".running_mean")
        put(s.bn_var, idx, bnkey + ".running_var")
    fw = np.ascontiguousarray(weights["fc.weight"], dtype=np.float32)
    fb = np.ascontiguousarray(weights["fc.bias"], dtype=np.float32)
    keep += [fw, fb]
    s.fc_w, s.fc_b = fw.ctypes.data, fb.ctypes.data
    for i in range(NUM_ACTS):
        s.act_scale[i] = float(act_scale[i])
    s.fp8 = 1 if fp8 else 0
    return s, keep


class ResNet18:
    """Whole-network INT8 runner (replaces main() of runtime/infer_e2e.cu for a batch)."""

    def __init__(self, ctx: Context, weights: Dict[str, np.ndarray], act_scale, max_batch: int, fp8: bool = False):
        """fp8=True: E4M3 weights / activations with FP32 accumulation (act_scale maps absmax to 448)"""
        self.ctx = ctx
        h = C.c_void_p()
        if isinstance(weights, WeightDir):       # loaded by the C++ loader; act_scale=None keeps the directory's scales
            if act_scale is not None:
                weights.set_act_scale(act_scale, fp8)
            ctx._ck(ctx.lib.dlq_resnet18_create(ctx.h, weights.struct_ptr, max_batch, C.byref(h)))
        else:
            s, keep = _weights_struct(weights, act_scale, fp8)
            ctx._ck(ctx.lib.dlq_resnet18_create(ctx.h, C.byref(s), max_batch, C.byref(h)))
        self.h = h
        self.max_batch = max_batch

    def forward(self, x, logits):
        self.ctx.order_after_torch()
        self.ctx._ck(self.ctx.lib.dlq_resnet18_forward(self.h, _ptr(x), x.shape[0], _ptr(logits)))

    @staticmethod
    def _hp(t):
        return t.data_ptr() if hasattr(t, "data_ptr") else t.ctypes.data

    def submit_host(self, x_host, logits_host):
        """pipelined form: enqueue H2D -> forward -> D2H and return (at most two in flight); pair with wait()"""
        self.ctx._ck(self.ctx.lib.dlq_resnet18_submit_host(self.h, self._hp(x_host), x_host.shape[0], self._hp(logits_host)))

    def submit_host_u8(self, x_hwc_host, logits_host):
        self.ctx._ck(self.ctx.lib.dlq_resnet18_submit_host_u8(self.h, self._hp(x_hwc_host), x_hwc_host.shape[0],
                                                              self._hp(logits_host)))

    def wait(self):
        """block until the oldest outstanding submit's logits are in host memory"""
        self.ctx._ck(self.ctx.lib.dlq_resnet18_wait(self.h))

    def launches_for_batch(self, n: int) -> int:
        return self.ctx.lib.dlq_resnet18_launches_for_batch(self.h, n)

    def set_option(self, key: str, value: int):
        self.ctx._ck(self.ctx.lib.dlq_resnet18_set_option(self.h, key.encode(), int(value)))

    def plan_info(self, n: int, key: str) -> int:
        v = C.c_int()
        self.ctx._ck(self.ctx.lib.dlq_resnet18_plan_info(self.h, n, key.encode(), C.byref(v)))
        return int(v.value)

    @property
    def dep_timeouts(self) -> int:
        n = C.c_uint()
        self.ctx._ck(self.ctx.lib.dlq_resnet18_dep_timeouts(self.h, C.byref(n)))
        return int(n.value)

    def enable_stamps(self, ring_forwards: int):
        self.ctx._ck(self.ctx.lib.dlq_resnet18_enable_stamps(self.h, ring_forwards))
        self._stamp_ring = ring_forwards

    def read_stamps(self) -> np.ndarray:
        """uint64 [forwards_recorded, launches, 2] = (first block entry, last block exit) in globaltimer ns"""
        out = np.zeros((self._stamp_ring, self.launches, 2), dtype=np.uint64)
        n = C.c_int()
        self.ctx._ck(self.ctx.lib.dlq_resnet18_read_stamps(self.h, out.ctypes.data, C.byref(n)))
        return out[:n.value]

    def forward_host(self, x_host, logits_host):
        """x_host / logits_host: CPU torch tensors (pinned for full speed) or numpy arrays."""
        xp = x_host.data_ptr() if hasattr(x_host, "data_ptr") else x_host.ctypes.data
        lp = logits_host.data_ptr() if hasattr(logits_host, "data_ptr") else logits_host.ctypes.data
        self.ctx._ck(self.ctx.lib.dlq_resnet18_forward_host(self.h, xp, x_host.shape[0], lp))

    IMAGENET_MEAN = (0.485, 0.456, 0.406)     # the reference's tools/preprocess_to_bin.py:5-6
    IMAGENET_STD = (0.229, 0.224, 0.225)

    def set_preprocess(self, mean=IMAGENET_MEAN, std=IMAGENET_STD):
        mean = np.asarray(mean, dtype=np.float32)
        std = np.asarray(std, dtype=np.float32)
        self.ctx._ck(self.ctx.lib.dlq_resnet18_set_preprocess(self.h, mean.ctypes.data, std.ctypes.data))

    def forward_u8(self, x_hwc, logits):
        """x_hwc: uint8 [N,224,224,3] device tensor (RGB); needs set_preprocess() first"""
        self.ctx.order_after_torch()
        self.ctx._ck(self.ctx.lib.dlq_resnet18_forward_u8(self.h, _ptr(x_hwc), x_hwc.shape[0], _ptr(logits)))

    def forward_host_u8(self, x_hwc_host, logits_host):
        xp = x_hwc_host.data_ptr() if hasattr(x_hwc_host, "data_ptr") else x_hwc_host.ctypes.data
        lp = logits_host.data_ptr() if hasattr(logits_host, "data_ptr") else logits_host.ctypes.data
        self.ctx._ck(self.ctx.lib.dlq_resnet18_forward_host_u8(self.h, xp, x_hwc_host.shape[0], lp))

    def checkpoint(self, name: str, out):
        self.ctx.order_after_torch()
        self.ctx._ck(self.ctx.lib.dlq_resnet18_checkpoint(self.h, name.encode(), _ptr(out)))

    def graph_capture(self, x, logits):
        """capture forward(x, logits) into a CUDA graph (x / logits must stay allocated); replay with graph_launch()"""
        self.ctx.order_after_torch()
        self.ctx._ck(self.ctx.lib.dlq_resnet18_graph_capture(self.h, _ptr(x), x.shape[0], _ptr(logits)))

    def graph_launch(self):
        self.ctx._ck(self.ctx.lib.dlq_resnet18_graph_launch(self.h))

    @property
    def launches(self) -> int:
        return self.ctx.lib.dlq_resnet18_launches(self.h)

    LAUNCH_NAMES = (["quantize_s2d", "conv1", "maxpool"]
                    + [n for b in range(8) for n in
                       ([f"layer{b // 2 + 1}.{b % 2}.conv1"]
                        + ([f"layer{b // 2 + 1}.{b % 2}.downsample"] if b in (2, 4, 6) else [])
                        + [f"layer{b // 2 + 1}.{b % 2}.conv2"])]
                    + ["gap_fc"])

    def profile(self, x, logits) -> np.ndarray:
        """per-launch device milliseconds of one forward (CUDA events on the context stream)"""
        ms = np.zeros(self.launches, dtype=np.float32)
        self.ctx.order_after_torch()
        self.ctx._ck(self.ctx.lib.dlq_resnet18_profile(self.h, _ptr(x), x.shape[0], _ptr(logits), ms.ctypes.data))
        return ms

    def close(self):
        if self.h:
            self.ctx.lib.dlq_resnet18_destroy(self.h)
            self.h = None


class ResNet18F32:
    """The reference's FP32 network on the GPU (bit-exact restatement of its operators): PTQ calibrator and the FP32
    side of the accuracy harness.  `weights` is a dict keyed like the reference's export, or a WeightDir."""

    CHECKPOINTS = {"stem_pool": (64, 56, 56), "layer1": (64, 56, 56), "layer2": (128, 28, 28), "layer3": (256, 14, 14),
                   "layer4": (512, 7, 7), "gap": (512,)}     # tools/diag_e2e_compare.py:5-13

    def __init__(self, ctx: Context, weights, max_batch: int):
        self.ctx = ctx
        h = C.c_void_p()
        if isinstance(weights, WeightDir):
            ctx._ck(ctx.lib.dlq_resnet18_f32_create(ctx.h, weights.struct_ptr, max_batch, C.byref(h)))
        else:
            s, keep = _weights_struct(weights, np.zeros(NUM_ACTS, np.float32))
            ctx._ck(ctx.lib.dlq_resnet18_f32_create(ctx.h, C.byref(s), max_batch, C.byref(h)))
        self.h = h
        self.max_batch = max_batch

    def forward(self, x, logits):
        self.ctx.order_after_torch()
        self.ctx._ck(self.ctx.lib.dlq_resnet18_f32_forward(self.h, _ptr(x), x.shape[0], _ptr(logits)))

    def checkpoint(self, name: str, out):
        self.ctx.order_after_torch()
        self.ctx._ck(self.ctx.lib.dlq_resnet18_f32_checkpoint(self.h, name.encode(), _ptr(out)))

    def absmax(self) -> np.ndarray:
        am = np.zeros(NUM_ACTS, dtype=np.float32)
        self.ctx._ck(self.ctx.lib.dlq_resnet18_f32_absmax(self.h, am.ctypes.data))
        return am

    def reset_absmax(self):
        self.ctx._ck(self.ctx.lib.dlq_resnet18_f32_reset_absmax(self.h))

    def act_scales(self, fp8: bool = False) -> np.ndarray:
        """PTQ activation scales from the running absmax (spec/QUANT_SPEC.md 2 / 6)"""
        am = self.absmax()
        sc = np.zeros(NUM_ACTS, dtype=np.float32)
        self.ctx.lib.dlq_act_scales_from_absmax(am.ctypes.data, 1 if fp8 else 0, sc.ctypes.data)
        return sc

    def close(self):
        if self.h:
            self.ctx.lib.dlq_resnet18_f32_destroy(self.h)
            self.h = None


class WeightDir:
    """A weight directory in the reference's export format (<key>.bin, tools/export_resnet18.py:85-92), read by the
    C++ loader.  No GPU needed."""

    def __init__(self, path: str):
        self.lib = load_library()
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = self.lib.dlq_weight_dir_load(path.encode(), C.byref(h), err, 512)
        if rc != 0:
            raise DlqError(rc, err.value.decode())
        self.h = h
        self.struct_ptr = self.lib.dlq_weight_dir_weights(h)

    @property
    def act_scale(self) -> np.ndarray:
        return np.array(list(self.struct_ptr.contents.act_scale), dtype=np.float32)

    def set_act_scale(self, act_scale, fp8: bool = False):
        for i in range(NUM_ACTS):
            self.struct_ptr.contents.act_scale[i] = float(act_scale[i])
        self.struct_ptr.contents.fp8 = 1 if fp8 else 0

    def tensor(self, field: str, idx: int, n: int) -> np.ndarray:
        """copy of one tensor, e.g. tensor('conv_w', 0, 64*3*7*7)"""
        p = getattr(self.struct_ptr.contents, field)
        p = p[idx] if idx is not None else p
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n,)).copy()

    def save(self, path: str, with_scales: bool = True):
        rc = self.lib.dlq_weight_dir_save(path.encode(), self.struct_ptr, 1 if with_scales else 0)
        if rc != 0:
            raise DlqError(rc, f"cannot write weight directory {path}")

    def close(self):
        if self.h:
            self.lib.dlq_weight_dir_free(self.h)
            self.h = None


def save_weight_dir(path: str, weights: Dict[str, np.ndarray], act_scale=None):
    """Write a dict keyed like the reference's export as <key>.bin + manifest.json through the C++ writer."""
    lib = load_library()
    s, keep = _weights_struct(weights, act_scale if act_scale is not None else np.zeros(NUM_ACTS, np.float32))
    os.makedirs(path, exist_ok=True)
    rc = lib.dlq_weight_dir_save(path.encode(), C.byref(s), 1 if act_scale is not None else 0)
    if rc != 0:
        raise DlqError(rc, f"cannot write weight directory {path}")


class MultiGPU:
    """Batch-sharded driver over several devices of one box (persistent worker thread per device).  The same device
    may be listed more than once (independent replicas)."""

    def __init__(self, devices, weights: Dict[str, np.ndarray], act_scale, max_batch_per_device: int, fp8: bool = False):
        self.lib = load_library()
        s, keep = _weights_struct(weights, act_scale, fp8)
        devs = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        rc = self.lib.dlq_multi_create(devs, len(devices), C.byref(s), max_batch_per_device, C.byref(h))
        if rc != 0:
            raise DlqError(rc, "dlq_multi_create failed")
        self.h = h
        self.n_devices = self.lib.dlq_multi_n_devices(h)

    @staticmethod
    def _hp(t):
        return t.data_ptr() if hasattr(t, "data_ptr") else t.ctypes.data

    def _ck(self, rc):
        if rc != 0:
            raise DlqError(rc, self.lib.dlq_multi_last_error_string(self.h).decode())

    def set_preprocess(self, mean=ResNet18.IMAGENET_MEAN, std=ResNet18.IMAGENET_STD):
        mean, std = np.asarray(mean, dtype=np.float32), np.asarray(std, dtype=np.float32)
        self._ck(self.lib.dlq_multi_set_preprocess(self.h, mean.ctypes.data, std.ctypes.data))

    def forward_host(self, x_host, logits_host):
        self._ck(self.lib.dlq_multi_forward_host(self.h, self._hp(x_host), x_host.shape[0], self._hp(logits_host)))

    def forward_host_u8(self, x_host, logits_host):
        self._ck(self.lib.dlq_multi_forward_host_u8(self.h, self._hp(x_host), x_host.shape[0], self._hp(logits_host)))

    def submit_host(self, x_host, logits_host):
        self._ck(self.lib.dlq_multi_submit_host(self.h, self._hp(x_host), x_host.shape[0], self._hp(logits_host)))

    def submit_host_u8(self, x_host, logits_host):
        self._ck(self.lib.dlq_multi_submit_host_u8(self.h, self._hp(x_host), x_host.shape[0], self._hp(logits_host)))

    def wait(self):
        self._ck(self.lib.dlq_multi_wait(self.h))

    def forward_device(self, xs, logits):
        """xs / logits: lists of device tensors, one per listed device (fp32 [n_g,3,224,224] / [n_g,1000])"""
        import torch
        for t in xs:
            torch.cuda.current_stream(t.device).synchronize()
        g = self.n_devices
        xp = (C.c_void_p * g)(*[t.data_ptr() for t in xs])
        lp = (C.c_void_p * g)(*[t.data_ptr() for t in logits])
        nn = (C.c_int * g)(*[int(t.shape[0]) for t in xs])
        self._ck(self.lib.dlq_multi_forward_device(self.h, xp, nn, lp))

    def close(self):
        if self.h:
            self.lib.dlq_multi_destroy(self.h)
            self.h = None


class MLP:
    """Two-layer MLP forward behind the C ABI (dlq_mlp_*): the GPU counterpart of the reference's MNIST forward
    (CUDA/MNIST_on_GPU/v4.cu:255-302, v5.cu:127-157, v3.c:177-215) on the tensor-core FC core.  w1 [in, hid] and
    w2 [hid, out] are in the reference's layout."""

    def __init__(self, ctx: Context, w1, b1, w2, b2, s_x: float, s_h: float, max_batch: int, fp8: bool = False):
        self.ctx = ctx
        w1 = np.ascontiguousarray(w1, dtype=np.float32)
        w2 = np.ascontiguousarray(w2, dtype=np.float32)
        b1 = np.ascontiguousarray(b1, dtype=np.float32)
        b2 = np.ascontiguousarray(b2, dtype=np.float32)
        self.n_in, self.n_hid = w1.shape
        self.n_out = w2.shape[1]
        h = C.c_void_p()
        ctx._ck(ctx.lib.dlq_mlp_create(ctx.h, w1.ctypes.data, b1.ctypes.data, w2.ctypes.data, b2.ctypes.data, self.n_in,
                                       self.n_hid, self.n_out, float(s_x), float(s_h), max_batch, 1 if fp8 else 0, C.byref(h)))
        self.h = h

    def forward(self, x, logits=None, probs=None):
        self.ctx.order_after_torch()
        self.ctx._ck(self.ctx.lib.dlq_mlp_forward(self.h, _ptr(x), x.shape[0], _ptr(logits), _ptr(probs)))

    def checkpoint(self, name: str, out):
        self.ctx.order_after_torch()
        self.ctx._ck(self.ctx.lib.dlq_mlp_checkpoint(self.h, name.encode(), _ptr(out)))

    def weight_scales(self, layer: int) -> np.ndarray:
        s = np.zeros(self.n_hid if layer == 1 else self.n_out, dtype=np.float32)
        self.ctx._ck(self.ctx.lib.dlq_mlp_weight_scales(self.h, layer, s.ctypes.data))
        return s

    def close(self):
        if self.h:
            self.ctx.lib.dlq_mlp_destroy(self.h)
            self.h = None
