"""GPU: the C++ drop-in driver examples/step8_e2e_dlq (the reference's `step8_e2e` command line on top of
libdlq_b200.so) - against the Python mirror of the same C ABI, and side by side with the REFERENCE's own
step8_e2e binary (oracle/_ref, built from runtime/infer_e2e.cu unmodified) on one weight directory and input file."""
import os
import re
import subprocess

import numpy as np
import pytest

import orc
from dlq_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRV = os.path.join(ROOT, "examples", "_build", "step8_e2e_dlq")
REF_E2E = os.path.join(orc.ORACLE_DIR, "_ref", "step8_e2e")
TOP1_RE = re.compile(r"\[E2E\]\s*top-1\s*class\s*index\s*=\s*(\d+)")     # tools/bench_fp32_vs_torch_e2e.py:30
NAMES = ["stem_pool", "layer1", "layer2", "layer3", "layer4", "gap", "logits"]


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    if not os.path.exists(DRV):
        pytest.skip("examples/_build/step8_e2e_dlq not built (make -C examples)")
    import dlq_b200
    d = tmp_path_factory.mktemp("drv")
    w = synth.make_weights(2)
    dlq_b200.save_weight_dir(str(d / "w"), w)
    synth.make_input(9, 1).tofile(str(d / "input1.bin"))
    synth.make_input(4, 5).tofile(str(d / "input5.bin"))
    synth.make_input(0, 8).tofile(str(d / "calib.bin"))
    return d, w


def _run(args, cwd):
    out = subprocess.run(args, capture_output=True, text=True, cwd=str(cwd), timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    return out.stdout


def test_driver_matches_python_mirror(ctx, workdir):
    """--calib + INT8 forward through the C++ driver == the same calls through the ctypes mirror, bit for bit"""
    import torch
    import dlq_b200
    d, w = workdir
    so = _run([DRV, "--manifest", str(d / "w"), "--input", str(d / "input5.bin"), "--calib", str(d / "calib.bin"),
               "--dump_dir", str(d / "dump5")], d)
    f = dlq_b200.ResNet18F32(ctx, w, 8)
    tmp = torch.empty((8, 1000), dtype=torch.float32, device="cuda")
    f.forward(torch.from_numpy(synth.make_input(0, 8)).cuda(), tmp)
    scales = f.act_scales()
    f.close()
    m = dlq_b200.ResNet18(ctx, w, scales, 5)
    dl = torch.empty((5, 1000), dtype=torch.float32, device="cuda")
    m.forward(torch.from_numpy(synth.make_input(4, 5)).cuda(), dl)
    ctx.sync()
    want = dl.cpu().numpy()
    got = np.fromfile(str(d / "dump5" / "logits.bin"), dtype=np.float32).reshape(5, 1000)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert [int(t) for t in TOP1_RE.findall(so)] == want.argmax(1).tolist()
    # dequantised checkpoints: int8 value * activation scale
    t8 = torch.empty((5, 512, 7, 7), dtype=torch.int8, device="cuda")
    m.checkpoint("layer4", t8)
    ctx.sync()
    l4 = np.fromfile(str(d / "dump5" / "layer4.bin"), dtype=np.float32).reshape(5, 512, 7, 7)
    assert np.array_equal(l4, t8.cpu().numpy().astype(np.float32) * np.float32(scales[4 + 3 * 7]))
    m.close()


def test_driver_side_by_side_with_reference_binary(workdir):
    """one weight directory, one input file, two binaries with the same command line: the reference's FP32 step8_e2e
    and ours.  --fp32 reproduces the reference's checkpoints (its own criterion: max_abs <= 1e-4, scaled); the INT8
    path agrees on top-1 and tracks the logits (cosine)."""
    if not os.path.exists(REF_E2E):
        pytest.skip("oracle/_ref/step8_e2e not built")
    d, _ = workdir
    args = ["--manifest", str(d / "w"), "--input", str(d / "input1.bin")]
    s_ref = _run([REF_E2E] + args + ["--dump_dir", str(d / "ref")], d)
    s_f32 = _run([DRV] + args + ["--dump_dir", str(d / "f32"), "--fp32"], d)
    s_i8 = _run([DRV] + args + ["--dump_dir", str(d / "i8"), "--calib", str(d / "calib.bin"), "--compare"], d)
    top_ref = int(TOP1_RE.search(s_ref).group(1))
    assert int(TOP1_RE.search(s_f32).group(1)) == top_ref
    assert int(TOP1_RE.search(s_i8).group(1)) == top_ref
    for name in NAMES:
        r = np.fromfile(str(d / "ref" / (name + ".bin")), dtype=np.float32)
        a = np.fromfile(str(d / "f32" / (name + ".bin")), dtype=np.float32)
        assert a.size == r.size
        assert float(np.abs(a - r).max()) / max(1.0, float(np.abs(r).max())) <= 1e-4, name
    r = np.fromfile(str(d / "ref" / "logits.bin"), dtype=np.float32).astype(np.float64)
    q = np.fromfile(str(d / "i8" / "logits.bin"), dtype=np.float32).astype(np.float64)
    assert float(r @ q / (np.linalg.norm(r) * np.linalg.norm(q))) > 0.99
    assert "agree_top1=1 (100.00%)" in s_i8 and "cosine=" in s_i8


def test_driver_parity_mode_exit_codes(workdir):
    """--expect_dir: the reference's test-binary convention (exit 0 = within atol, exit 2 = mismatch).  The FP32
    arithmetic passes against the reference binary's own dump at the reference's 1e-4; the INT8 path, compared at
    that FP32 criterion, is reported as a mismatch - and passes at a quantisation-sized tolerance."""
    if not os.path.exists(REF_E2E):
        pytest.skip("oracle/_ref/step8_e2e not built")
    d, _ = workdir
    args = ["--manifest", str(d / "w"), "--input", str(d / "input1.bin")]
    _run([REF_E2E] + args + ["--dump_dir", str(d / "refp")], d)
    r = subprocess.run([DRV] + args + ["--fp32", "--dump_dir", str(d / "p32"), "--expect_dir", str(d / "refp")],
                       capture_output=True, text=True, cwd=str(d), timeout=300)
    assert r.returncode == 0 and r.stdout.count("[OK]") == 7 and "[FAIL]" not in r.stdout, r.stdout + r.stderr
    q = [DRV] + args + ["--calib", str(d / "calib.bin"), "--dump_dir", str(d / "p8"), "--expect_dir", str(d / "refp")]
    r = subprocess.run(q, capture_output=True, text=True, cwd=str(d), timeout=300)
    assert r.returncode == 2 and "[FAIL]" in r.stdout, r.stdout + r.stderr
    r = subprocess.run(q + ["--atol", "0.1"], capture_output=True, text=True, cwd=str(d), timeout=300)
    assert r.returncode == 0 and r.stdout.count("[OK]") == 7, r.stdout + r.stderr


def test_driver_usage_and_io_errors(workdir):
    d, _ = workdir
    assert subprocess.run([DRV], capture_output=True).returncode == 1                       # usage -> 1 (R/infer_e2e.cu:241)
    r = subprocess.run([DRV, "--manifest", str(d / "nope"), "--input", str(d / "input1.bin")], capture_output=True, text=True)
    assert r.returncode == 1 and "open fail" in r.stderr
    (d / "short.bin").write_bytes(b"\0" * 4000)
    r = subprocess.run([DRV, "--manifest", str(d / "w"), "--input", str(d / "short.bin")], capture_output=True, text=True)
    assert r.returncode == 1 and "unexpected size" in r.stderr


def test_cpp_bench_driver(ctx):
    """examples/dlq_bench: synthetic weights generated and PTQ-calibrated in C++ (the committed scales, bit for bit),
    then timed forwards; its top-1 for image 0 equals the ctypes mirror's on the same synthetic image"""
    import json
    import torch
    import dlq_b200
    exe = os.path.join(ROOT, "examples", "_build", "dlq_bench")
    if not os.path.exists(exe):
        pytest.skip("examples/_build/dlq_bench not built (make -C examples)")
    out = subprocess.run([exe, "--batch", "8", "--iters", "3", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["n_gpus"] == 1 and d["batch"] == 8 and d["images_per_s"] > 0 and d["host_images_per_s"] > 0
    assert 50 < d["latency_b1_us"] < 5000 and d["launches_per_forward"] == 20      # batch 8: shortcut convs fused
    assert d["host_u8_pipelined_images_per_s"] > 0
    w = synth.make_weights(0)
    m = dlq_b200.ResNet18(ctx, w, synth.load_act_scales(0), 8)
    dl = torch.empty((8, 1000), dtype=torch.float32, device="cuda")
    m.forward(torch.from_numpy(synth.make_input(0, 8)).cuda(), dl)
    ctx.sync()
    assert d["top1_image0"] == int(dl[0].argmax().item())
    m.close()
    # the batch-sharded driver: two replicas on device 0 (runs on a one-GPU box), and two devices when the box has them
    lists = [["--devices", "0,0"]] + ([["--gpus", "2"]] if torch.cuda.device_count() >= 2 else [])
    for extra in lists:
        out = subprocess.run([exe, "--batch", "16", "--iters", "2", "--warmup", "1"] + extra, capture_output=True,
                             text=True, timeout=300)
        assert out.returncode == 0, out.stdout + out.stderr
        d2 = json.loads(out.stdout.strip().splitlines()[-1])
        assert d2["n_gpus"] == 2 and d2["batch_per_gpu"] == 8 and d2["top1_image0"] == d["top1_image0"]
        for k in ("host_images_per_s", "host_pipelined_images_per_s", "host_u8_pipelined_images_per_s",
                  "device_resident_images_per_s"):
            assert d2[k] > 0, k
    assert subprocess.run([exe, "--bogus"], capture_output=True).returncode == 1
