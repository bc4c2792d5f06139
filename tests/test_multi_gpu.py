"""GPU: the batch-sharded multi-GPU driver (dlq_multi_*, north_star subsystem 3) - logits bit-equal to a single
context's, for every entry point: fp32 / uint8 host buffers, submit + wait, device-resident.  Runs on a ONE-GPU box by
listing device 0 twice (two replicas, two worker threads, separate streams); the real two-device variant runs when the
box has two GPUs.  Reference wiring: runtime/infer_e2e.cu:254-433 (one image, one device)."""
import numpy as np
import pytest

from dlq_b200 import synth

pytestmark = pytest.mark.gpu


def _single_context_logits(x, u8=None):
    import torch
    import dlq_b200
    ctx = dlq_b200.Context(0)
    m = dlq_b200.ResNet18(ctx, synth.make_weights(0), synth.load_act_scales(0), x.shape[0])
    dl = torch.empty((x.shape[0], 1000), dtype=torch.float32, device="cuda")
    m.forward(torch.from_numpy(x).cuda(), dl)
    ctx.sync()
    out = dl.cpu().numpy()
    out_u8 = None
    if u8 is not None:
        m.set_preprocess()
        du = torch.empty((u8.shape[0], 1000), dtype=torch.float32, device="cuda")
        m.forward_u8(torch.from_numpy(u8).cuda(), du)
        ctx.sync()
        out_u8 = du.cpu().numpy()
    m.close()
    ctx.close()
    return out, out_u8


def _device_lists():
    import torch
    lists = [[0, 0], [0, 0, 0]]
    if torch.cuda.device_count() >= 2:
        lists.append([0, 1])
    return lists


def test_multi_logits_equal_single_context():
    import torch
    import dlq_b200
    n = 21                                    # ragged split: 11 + 10, or 7 + 7 + 7
    x = synth.make_input(31, n)
    rng = np.random.default_rng(3)
    u8 = rng.integers(0, 256, (n, 224, 224, 3), dtype=np.uint8)
    want, want_u8 = _single_context_logits(x, u8)
    w, sc = synth.make_weights(0), synth.load_act_scales(0)
    for devs in _device_lists():
        mg = dlq_b200.MultiGPU(devs, w, sc, 16)
        assert mg.n_devices == len(devs)
        xh = torch.from_numpy(x).pin_memory()
        lh = torch.zeros((n, 1000), dtype=torch.float32).pin_memory()
        mg.forward_host(xh, lh)
        assert np.array_equal(lh.numpy().view(np.uint32), want.view(np.uint32)), devs
        # uint8 entry
        mg.set_preprocess()
        uh = torch.from_numpy(u8).pin_memory()
        lh.zero_()
        mg.forward_host_u8(uh, lh)
        assert np.array_equal(lh.numpy().view(np.uint32), want_u8.view(np.uint32)), devs
        # asynchronous: three batches in flight before one wait
        outs = [torch.zeros((n, 1000), dtype=torch.float32).pin_memory() for _ in range(3)]
        mg.submit_host(xh, outs[0])
        mg.submit_host_u8(uh, outs[1])
        mg.submit_host(xh[:5], outs[2][:5])
        mg.wait()
        assert np.array_equal(outs[0].numpy().view(np.uint32), want.view(np.uint32))
        assert np.array_equal(outs[1].numpy().view(np.uint32), want_u8.view(np.uint32))
        assert np.array_equal(outs[2].numpy()[:5].view(np.uint32), want[:5].view(np.uint32))
        # device-resident: every replica gets its own shard
        g = len(devs)
        per = (n + g - 1) // g
        xs, ls = [], []
        for i, d in enumerate(devs):
            lo, hi = min(n, i * per), min(n, (i + 1) * per)
            xs.append(torch.from_numpy(x[lo:hi]).to(f"cuda:{d}"))
            ls.append(torch.empty((hi - lo, 1000), dtype=torch.float32, device=f"cuda:{d}"))
        mg.forward_device(xs, ls)
        got = np.concatenate([t.cpu().numpy() for t in ls], 0)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), devs
        # a batch larger than the replicas hold is an argument error, not a crash
        big = torch.zeros((16 * g + 1, 3, 224, 224), dtype=torch.float32)
        with pytest.raises(dlq_b200.DlqError) as ei:
            mg.forward_host(big, torch.zeros((16 * g + 1, 1000), dtype=torch.float32))
        assert ei.value.code == 1
        mg.close()


def test_multi_fp8_replicas_agree():
    """the E4M3 network through the driver: replicas produce the same bytes as one context (single issuer per
    accumulator: no timing dependence)"""
    import torch
    import dlq_b200
    n = 6
    x = synth.make_input(2, n)
    w = synth.make_weights(0)
    s8 = (np.asarray(synth.load_act_scales(0), dtype=np.float64) * 127.0 / 448.0).astype(np.float32)
    ctx = dlq_b200.Context(0)
    m = dlq_b200.ResNet18(ctx, w, s8, n, fp8=True)
    dl = torch.empty((n, 1000), dtype=torch.float32, device="cuda")
    m.forward(torch.from_numpy(x).cuda(), dl)
    ctx.sync()
    want = dl.cpu().numpy()
    m.close()
    ctx.close()
    mg = dlq_b200.MultiGPU([0, 0], w, s8, n, fp8=True)
    lh = torch.zeros((n, 1000), dtype=torch.float32).pin_memory()
    mg.forward_host(torch.from_numpy(x).pin_memory(), lh)
    assert np.array_equal(lh.numpy().view(np.uint32), want.view(np.uint32))
    mg.close()
