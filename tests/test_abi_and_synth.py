"""CPU: the C-ABI library loads and exports every symbol include/dlq.h declares (no compute calls without a GPU);
the product's synthetic generator equals the oracle's; host-side folding helpers equal the oracle's; the batch
sharding used by the multi-GPU driver / bench.py is a partition."""
import ctypes
import os
import re

import numpy as np
import pytest

import orc
import dlq_b200
from dlq_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "dlq.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dlq_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = dlq_b200.load_library()
    syms = _header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/dlq.h but not exported by libdlq_b200.so"
    assert sorted(dlq_b200.ABI_SYMBOLS) == syms, set(dlq_b200.ABI_SYMBOLS) ^ set(syms)
    assert b"sm_100a" in lib.dlq_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(dlq_b200.DlqError):
        dlq_b200.Context(0)


def test_synth_generators_agree():
    for name, lo, hi, sh in [("conv1.weight", -127, 127, 9), ("bn1.weight", 128, 384, 8), ("input", -192, 192, 6)]:
        a = dlq_b200.synth_fill_f32((1000,), 2, name, lo, hi, sh)
        b = orc.fill_f32((1000,), 2, name, lo, hi, sh)
        assert np.array_equal(a, b)
    wa, wb = synth.make_weights(1), synth.make_weights(1, fill=orc.fill_f32)
    assert wa.keys() == wb.keys() and all(np.array_equal(wa[k], wb[k]) for k in wa)
    assert len(wa) == 20 * 5 + 2 and wa["layer4.1.conv2.weight"].shape == (512, 512, 3, 3)
    assert synth.load_act_scales(0).shape == (27,)


def test_fold_bn_matches_oracle():
    rng = np.random.default_rng(0)
    g, b, m = (rng.standard_normal(64).astype(np.float32) for _ in range(3))
    v = rng.uniform(0.5, 1.5, 64).astype(np.float32)
    sw = rng.uniform(1e-3, 1e-2, 64).astype(np.float32)
    a1, b1 = dlq_b200.fold_bn(g, b, m, v, sw, 0.02, 0.3)
    a2, b2 = orc.fold_bn(g, b, m, v, sw, 0.02, 0.3)
    assert np.array_equal(a1, a2) and np.array_equal(b1, b2)
    assert dlq_b200.res_mul(0.37, 0.11) == orc.res_mul(0.37, 0.11)


def test_batch_sharding_is_a_partition():
    import bench
    for total, world in [(2048, 8), (256, 1), (10, 4), (7, 8), (0, 2)]:
        spans = [bench.shard(total, world, r) for r in range(world)]
        covered = [i for lo, hi in spans for i in range(lo, hi)]
        assert covered == list(range(total))
