"""Weight-directory loader / writer on the CPU: the reference's export layout
(tools/export_resnet18.py:85-92, names as runtime/infer_e2e.cu:262-330,428-429 reads them) and load_bin_f32's
size checks (runtime/utils.hpp:48-60), reported as errors instead of exit(1)."""
import json
import os

import numpy as np
import pytest

import dlq_b200
from dlq_b200 import synth


@pytest.fixture(scope="module")
def wdir(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("weights"))
    w = synth.make_weights(0)
    dlq_b200.save_weight_dir(d, w, synth.load_act_scales(0))
    return d, w


def test_layout_matches_reference_export(wdir):
    d, w = wdir
    files = set(os.listdir(d))
    # every key the reference driver loads (runtime/infer_e2e.cu:262-330,428-429)
    for key in ["conv1.weight", "bn1.weight", "bn1.bias", "bn1.running_mean", "bn1.running_var", "fc.weight", "fc.bias",
                "layer1.0.conv1.weight", "layer2.0.downsample.0.weight", "layer2.0.downsample.1.running_var",
                "layer4.1.bn2.bias"]:
        assert key + ".bin" in files
        assert np.array_equal(np.fromfile(os.path.join(d, key + ".bin"), dtype="<f4"), w[key].ravel())
    man = json.load(open(os.path.join(d, "manifest.json")))
    assert man["model"] == "resnet18" and man["dtype"] == "fp32" and man["layout"] == "NCHW"
    assert man["preprocess"]["mean"] == [0.485, 0.456, 0.406]
    assert man["tensors"]["conv1.weight"] == {"shape": [64, 3, 7, 7], "layout": "OIHW", "kind": "conv_weight",
                                              "path": "conv1.weight.bin"}
    assert man["tensors"]["fc.weight"]["layout"] == "OI" and man["tensors"]["bn1.running_mean"]["kind"] == "bn_buffer"
    assert len(man["tensors"]) == 102 and man["quant"]["num_act_scales"] == 27


def test_loader_round_trip(wdir):
    d, w = wdir
    wd = dlq_b200.WeightDir(d)
    assert np.array_equal(wd.act_scale, np.asarray(synth.load_act_scales(0), np.float32))
    assert np.array_equal(wd.tensor("conv_w", 0, 64 * 3 * 49), w["conv1.weight"].ravel())
    assert np.array_equal(wd.tensor("conv_w", 23, 512 * 512 * 9), w["layer4.1.conv2.weight"].ravel())
    assert np.array_equal(wd.tensor("bn_var", 9, 128), w["layer2.0.downsample.1.running_var"].ravel())
    assert np.array_equal(wd.tensor("fc_b", None, 1000), w["fc.bias"].ravel())
    assert not wd.struct_ptr.contents.conv_w[3]          # layer1.0 has no downsample conv
    wd.close()


def test_loader_errors(wdir, tmp_path):
    d, _ = wdir
    import shutil
    bad = str(tmp_path / "bad")
    shutil.copytree(d, bad)
    os.remove(os.path.join(bad, "layer3.0.downsample.1.bias.bin"))
    with pytest.raises(dlq_b200.DlqError, match="open fail"):
        dlq_b200.WeightDir(bad)
    np.zeros(255, dtype="<f4").tofile(os.path.join(bad, "layer3.0.downsample.1.bias.bin"))
    with pytest.raises(dlq_b200.DlqError, match="unexpected size"):
        dlq_b200.WeightDir(bad)
    with open(os.path.join(bad, "layer3.0.downsample.1.bias.bin"), "wb") as f:
        f.write(b"\0" * 1023)
    with pytest.raises(dlq_b200.DlqError, match="not float-aligned"):
        dlq_b200.WeightDir(bad)


def test_act_scales_from_absmax():
    lib = dlq_b200.load_library()
    am = np.array([3.0, 0.0] + [float(i) for i in range(25)], dtype=np.float32)
    s = np.zeros(27, dtype=np.float32)
    lib.dlq_act_scales_from_absmax(am.ctypes.data, 0, s.ctypes.data)
    want = (np.where(am > 0, am, np.float32(1)).astype(np.float64) / 127.0).astype(np.float32)
    assert np.array_equal(s, want)
    lib.dlq_act_scales_from_absmax(am.ctypes.data, 1, s.ctypes.data)
    assert np.array_equal(s, (np.where(am > 0, am, np.float32(1)).astype(np.float64) / 448.0).astype(np.float32))
