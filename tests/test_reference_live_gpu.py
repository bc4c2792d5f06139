"""GPU: runs the REFERENCE's own FP32 CUDA kernels (compiled unmodified into oracle/_ref by oracle/Makefile,
which travels to the GPU box prebuilt) and checks the CPU oracle's FP32 restatement against them — the live pin
SURVEY §8c asks for.  Skipped when oracle/_ref has not been built."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

import orc
from dlq_b200 import synth

pytestmark = pytest.mark.gpu
REF = os.path.join(orc.ORACLE_DIR, "_ref")
LIB = os.path.join(REF, "libref_rkl_fp32.so")
E2E = os.path.join(REF, "step8_e2e")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(LIB):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return C.CDLL(LIB)


def _d(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_reference_conv_bn_relu_kernels_match_oracle(ref):
    import torch
    x = orc.fill_f32((1, 16, 20, 20), 0, "ref.x", -192, 192, 6)
    w = orc.fill_f32((24, 16, 3, 3), 0, "ref.w", -127, 127, 9)
    K, OH = 16 * 9, 20
    dx, dw = _d(x), _d(w.reshape(24, K))
    col = torch.zeros((K, OH * OH), dtype=torch.float32, device="cuda")
    y = torch.zeros((24, OH * OH), dtype=torch.float32, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    assert ref.ref_im2col_nchw(p(dx), 1, 16, 20, 20, 3, 3, 1, 1, 1, 1, p(col)) == 0
    assert ref.ref_sgemm_tiled(p(dw), p(col), p(y), 24, OH * OH, K) == 0
    want = orc.conv2d_f32(x, w, 1, 1).reshape(24, OH * OH)
    # sequential-K FMA chain on both sides (K/sgemm_tiled.cu:22-40) -> identical bits
    assert np.array_equal(y.cpu().numpy(), want)
    g, b, m = (orc.fill_f32((24,), 0, n, -64, 64, 8) for n in ("g", "b", "m"))
    v = orc.fill_f32((24,), 0, "v", 128, 384, 8)
    assert ref.ref_bn_inference(p(y), p(_d(g)), p(_d(b)), p(_d(m)), p(_d(v)), C.c_float(1e-5), 24, OH, OH) == 0
    want = orc.bn_inference_f32(want.reshape(1, 24, OH, OH), g, b, m, v)
    assert np.abs(y.cpu().numpy().reshape(want.shape) - want).max() <= 1e-5     # nvcc div/sqrt are IEEE; FMA contraction may differ
    assert ref.ref_relu_forward(p(y), 24 * OH * OH) == 0
    assert np.abs(y.cpu().numpy().reshape(want.shape) - orc.relu_f32(want)).max() <= 1e-5


def test_reference_pool_gap_kernels_match_oracle(ref):
    import torch
    x = orc.fill_f32((2, 8, 14, 14), 1, "ref.p", -192, 192, 6)
    p = lambda t: C.c_void_p(t.data_ptr())
    y = torch.zeros((2, 8, 7, 7), dtype=torch.float32, device="cuda")
    assert ref.ref_maxpool2d_3x3_s2p1_nchw(p(_d(x)), 2, 8, 14, 14, p(y)) == 0
    assert np.array_equal(y.cpu().numpy(), orc.maxpool_f32(x))
    gy = torch.zeros(8, dtype=torch.float32, device="cuda")
    assert ref.ref_gap_global(p(_d(x[0])), 8, 14, 14, p(gy)) == 0
    assert np.array_equal(gy.cpu().numpy(), orc.gap_f32(x[:1])[0])          # same strided sums + smem tree order


def test_reference_step8_e2e_binary_matches_oracle(tmp_path):
    """The reference's whole-network driver (R/infer_e2e.cu, unmodified) on synthetic weights written in its own
    <key>.bin format; its --dump_dir checkpoints vs the FP32 oracle (reference criterion atol 1e-4, scaled)."""
    if not os.path.exists(E2E):
        pytest.skip("oracle/_ref/step8_e2e not built")
    w = synth.make_weights(2, fill=orc.fill_f32)
    wdir, ddir = tmp_path / "w", tmp_path / "dump"
    wdir.mkdir()
    for k, a in w.items():
        a.astype(np.float32).tofile(str(wdir / (k + ".bin")))
    x = synth.make_input(9, 1, fill=orc.fill_f32)
    x.tofile(str(tmp_path / "input.bin"))
    out = subprocess.run([E2E, "--manifest", str(wdir), "--input", str(tmp_path / "input.bin"), "--dump_dir", str(ddir)],
                         capture_output=True, text=True, cwd=str(tmp_path), timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    got = orc.F32Model(w).forward(x, checkpoints=True)
    shapes = {"stem_pool": (1, 64, 56, 56), "layer1": (1, 64, 56, 56), "layer2": (1, 128, 28, 28),
              "layer3": (1, 256, 14, 14), "layer4": (1, 512, 7, 7), "gap": (1, 512), "logits": (1, 1000)}
    for name, shp in shapes.items():
        ref = np.fromfile(str(ddir / (name + ".bin")), dtype=np.float32).reshape(shp)
        scale = max(1.0, float(np.abs(ref).max()))
        err = float(np.abs(got[name] - ref).max()) / scale
        assert err <= 1e-4, (name, err)
    top = int(got["logits"].argmax())
    assert f"top-1 class index = {top}" in out.stdout
