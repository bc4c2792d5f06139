"""GPU parity: dlq_conv2d_i8 (tcgen05 implicit GEMM + fused epilogue) vs the CPU oracle, through the C ABI.
Bar: bit-exact int32 accumulators and int8 outputs (QUANT_SPEC §3)."""
import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu

# (name, N, IC, H, W, OC, k, stride, pad)
SHAPES = [
    ("layer1.conv", 2, 64, 56, 56, 64, 3, 1, 1),
    ("layer2.0.conv1", 2, 64, 56, 56, 128, 3, 2, 1),
    ("layer2.0.downsample", 2, 64, 56, 56, 128, 1, 2, 0),
    ("layer2.conv", 3, 128, 28, 28, 128, 3, 1, 1),
    ("layer3.0.conv1", 2, 128, 28, 28, 256, 3, 2, 1),
    ("layer3.0.downsample", 2, 128, 28, 28, 256, 1, 2, 0),
    ("layer3.conv", 3, 256, 14, 14, 256, 3, 1, 1),
    ("layer4.0.conv1", 2, 256, 14, 14, 512, 3, 2, 1),
    ("layer4.0.downsample", 5, 256, 14, 14, 512, 1, 2, 0),
    ("layer4.conv", 5, 512, 7, 7, 512, 3, 1, 1),
    ("conv1.stem", 2, 3, 224, 224, 64, 7, 2, 3),
    ("ragged.1x1.s1", 1, 64, 5, 9, 64, 1, 1, 0),
    ("ragged.3x3.odd", 1, 128, 9, 11, 64, 3, 1, 1),
    ("ragged.5x5", 2, 64, 12, 10, 64, 5, 1, 2),
]


@pytest.mark.parametrize("shape", SHAPES, ids=[s[0] for s in SHAPES])
@pytest.mark.parametrize("seed", [0, 1])
def test_conv_parity(ctx, shape, seed):
    import torch
    name, N, IC, H, W, OC, k, stride, pad = shape
    rng = np.random.default_rng(1000 * seed + len(name))
    x = orc.fill_i8((N, IC, H, W), seed, name + ".x")
    wq = orc.fill_i8((OC, IC, k, k), seed, name + ".w", -127, 127)
    OH, OW = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    alpha = (rng.uniform(0.5, 1.5, OC) * 2.0 ** -8).astype(np.float32)
    beta = rng.uniform(-40, 40, OC).astype(np.float32)
    res = orc.fill_i8((N, OC, OH, OW), seed, name + ".res")
    res_mul = np.float32(0.74)

    w = ctx.pack_conv_weights_i8(wq, stride, pad)
    dx = torch.from_numpy(x).cuda()
    dal, dbe = torch.from_numpy(alpha).cuda(), torch.from_numpy(beta).cuda()
    dres = torch.from_numpy(res).cuda()
    for use_res, relu in [(True, True), (False, False)]:
        dy = torch.full((N, OC, OH, OW), 77, dtype=torch.int8, device="cuda")
        dacc = torch.full((N, OC, OH, OW), -12345, dtype=torch.int32, device="cuda")
        oh, ow = ctx.conv2d_i8(dx, w, dal, dbe, dres if use_res else None, float(res_mul), relu, dy, dacc)
        ctx.sync()
        assert (oh, ow) == (OH, OW)
        acc_ref, y_ref = orc.conv2d_i8(x, wq, stride, pad, alpha, beta, res if use_res else None, float(res_mul), relu)
        acc = dacc.cpu().numpy()
        y = dy.cpu().numpy()
        bad_acc = int((acc != acc_ref).sum())
        assert bad_acc == 0, f"{name}: {bad_acc}/{acc.size} int32 accumulators differ"
        bad_y = int((y != y_ref).sum())
        assert bad_y == 0, f"{name}: {bad_y}/{y.size} int8 outputs differ (res={use_res}, relu={relu})"
    w.free()


def test_conv_bad_arguments(ctx):
    import dlq_b200
    with pytest.raises(dlq_b200.DlqError) as e:
        ctx.pack_conv_weights_i8(np.zeros((64, 48, 3, 3), np.int8), 1, 1)   # IC not 64 / multiple of 128
    assert e.value.code == 1
    with pytest.raises(dlq_b200.DlqError) as e:
        ctx.pack_conv_weights_i8(np.zeros((64, 64, 3, 3), np.int8), 3, 1)   # stride 3
    assert e.value.code == 1
