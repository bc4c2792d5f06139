#!/usr/bin/env python
"""bench.py — ResNet-18 INT8 throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

N=1: plain python.  N>1: launched by torch.distributed.run, one rank per GPU; images are batch-sharded
(weak scaling: every rank processes `batch` images per step), no data-path collective — the only
cross-rank operations are the timing barrier and the max-reduction of the elapsed time.

A "step" is one pass of the hot path (fp32 NCHW image batch -> int8 quantise -> 20 tcgen05 convs with fused
epilogue -> max-pool -> GAP+FC -> fp32 logits) over one batch of 256 synthetic images.
  value   images/s with the input batch resident in HBM (CUDA events on the library's stream)
  e2e     the same through dlq_resnet18_forward_host: pinned HOST input, H2D copy, forward, D2H of the logits
  roofline  the conv kernel family (conv_i8_kernel<ROWB>, 20 launches/step): algorithmic int8 ops
            (3.627 GOP/img) / the time the conv launches occupy the GPU inside the running step (union of
            their [first block entry, last block exit] globaltimer spans), against 2 x measured bf16 dense - the burst
            figure when the SM clock sampled during the run is >= 1.9 GHz, else the sustained one
  cpu_baseline  the CPU oracle (restatement + QUANT_SPEC; the reference has no CPU ResNet path) on the
            box's host cores over a bounded sample

--impl reference: the reference arm.  The reference repo contains no CPU (or INT8) implementation of this
path, so the arm times the CPU oracle port with all host threads on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONV_GOP_PER_IMG = 3.62712        # 2 x 1813.56 MMAC
CONV_BYTES_PER_IMG = 5462736      # algorithmic conv bytes/img at B=256
METRIC = "resnet18_int8_images_per_sec"
WORKLOAD = "ResNet-18 INT8 (per-channel weight / per-tensor activation), batch 256, 224x224, fused conv epilogue"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.p = None
        self.thread = None
        self.samples = []
        self.stop_flag = False

    # NVML in-process sampling (every ~2 ms: the timed region of a short run lasts tens of milliseconds, too short
    # for `nvidia-smi -lms`); the nvidia-smi subprocess is the fallback
    def _nvml_loop(self, handle, nv):
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(handle)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                self.samples.append((float(sm), int(rs)))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.nv = nv
            self.stop_flag = False
            self.samples = []
            self.thread = threading.Thread(target=self._nvml_loop, args=(h, nv), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(self.index), "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            if self.samples:
                bits = 0
                for _, r in self.samples:
                    bits |= r
                names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
                return {"sm_mhz": statistics.median([c for c, _ in self.samples]), "sm_max_mhz": self.max_mhz,
                        "reasons": sorted(n for b, n in names.items() if bits & b), "samples": len(self.samples),
                        "how": "NVML, sampled every 2 ms during the timed region"}
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.splitlines():
            f = [s.strip() for s in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def shard(total: int, world: int, rank: int):
    """contiguous batch split used by the multi-GPU driver: rank r takes [r*ceil(total/world), ...)"""
    per = (total + world - 1) // world
    lo = min(total, rank * per)
    return lo, min(total, lo + per)


def cpu_oracle_rate(sample_images: int):
    """images/s of the CPU oracle INT8 forward on `sample_images` synthetic images (all host threads)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    from dlq_b200 import synth
    orc.lib().orc_set_num_threads(os.cpu_count() or 1)
    w = synth.make_weights(0, fill=orc.fill_f32)
    m = orc.I8Model(w, synth.load_act_scales(0))
    x = synth.make_input(0, sample_images, fill=orc.fill_f32)
    t0 = time.perf_counter()
    m.forward(x)
    dt = time.perf_counter() - t0
    return sample_images / dt, orc.lib().orc_num_threads(), dt


def mnist_config0(ctx=None):
    """BASELINE configs[0]: MNIST MLP FP32 forward, batch 1024, on the reference's host C path: MN/v3.c forward_timed
    from oracle/_ref/libref_mnist_v3.so when that was built (kind "reference"), else the oracle restatement (kind
    "port"); single thread, as the reference is.  With a context, also the INT8 MLP on the GPU (dlq_b200/mnist.py)."""
    import ctypes as C
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    B, ind, hid, outd = 1024, 784, 256, 10
    rng = np.random.default_rng(12345)
    x = ((rng.random((B, ind), dtype=np.float32) - np.float32(0.1307)) / np.float32(0.3081)).astype(np.float32)
    w1 = (rng.standard_normal((ind, hid)) * np.sqrt(2.0 / ind)).astype(np.float32)
    w2 = (rng.standard_normal((hid, outd)) * np.sqrt(2.0 / hid)).astype(np.float32)
    b1, b2 = np.zeros(hid, np.float32), np.zeros(outd, np.float32)
    out = {"workload": "MNIST MLP 784-256-10 forward, batch 1024, synthetic data (MN/v3.c:99-105 normalisation)"}
    ref = os.path.join(ROOT, "oracle", "_ref", "libref_mnist_v3.so")
    hidden, y = np.zeros((B, hid), np.float32), np.zeros((B, outd), np.float32)
    fp = C.POINTER(C.c_float)
    reps = 3
    if os.path.exists(ref):
        lib = C.CDLL(ref)

        class NN(C.Structure):
            _fields_ = [(n, fp) for n in ("weights1", "weights2", "bias1", "bias2", "grad_weights1", "grad_weights2",
                                          "grad_bias1", "grad_bias2")]
        nn = NN()
        nn.weights1, nn.weights2 = w1.ctypes.data_as(fp), w2.ctypes.data_as(fp)
        nn.bias1, nn.bias2 = b1.ctypes.data_as(fp), b2.ctypes.data_as(fp)
        stats = (C.c_double * 32)()
        t0 = time.perf_counter()
        for _ in range(reps):
            lib.forward_timed(C.byref(nn), x.ctypes.data_as(fp), hidden.ctypes.data_as(fp), y.ctypes.data_as(fp), B, stats)
        dt = (time.perf_counter() - t0) / reps
        out["cpu"] = {"kind": "reference", "impl": "MN/v3.c forward_timed (unmodified, gcc -O2)", "cores": 1,
                      "ms_per_forward": dt * 1e3, "images_per_s": B / dt}
    else:
        import orc
        t0 = time.perf_counter()
        for _ in range(reps):
            orc.mnist_forward(x, w1, b1, w2, b2)
        dt = (time.perf_counter() - t0) / reps
        out["cpu"] = {"kind": "port", "impl": "oracle restatement of MN/v3.c:125-215", "cores": 1,
                      "ms_per_forward": dt * 1e3, "images_per_s": B / dt}
    if ctx is not None:
        import torch
        from dlq_b200.mnist import MnistMLP
        m = MnistMLP(ctx, w1, b1, w2, b2, x[:128], max_batch=B)
        dx = torch.from_numpy(x).cuda()
        dz = torch.empty((B, outd), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        stream = torch.cuda.ExternalStream(ctx.stream)
        for _ in range(5):
            m.forward(dx, dz, None)
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(50):
            m.forward(dx, dz, None)
        e1.record(stream)
        ctx.sync()
        ms = e0.elapsed_time(e1) / 50
        out["gpu_int8"] = {"ms_per_forward": ms, "images_per_s": B / (ms * 1e-3),
                           "how": "dlq_mlp_forward (C ABI): quantise, two FC layers on the tcgen05 GEMM core, head - 4 launches, "
                                  "device-resident input"}
        m.close()
    return out


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    from dlq_b200 import synth
    w = synth.make_weights(0, fill=orc.fill_f32)
    m = orc.I8Model(w, synth.load_act_scales(0))
    orc.lib().orc_set_num_threads(os.cpu_count() or 1)     # torchrun exports OMP_NUM_THREADS=1
    threads = orc.lib().orc_num_threads()
    x1 = synth.make_input(0, threads, fill=orc.fill_f32)
    t0 = time.perf_counter()
    m.forward(x1)
    t_probe = (time.perf_counter() - t0) / threads      # seconds per image with all threads busy
    budget = min(2.0, 150.0 / max(1, args.steps + args.warmup))
    per_step = max(threads, int(budget / t_probe) // threads * threads)
    x = synth.make_input(0, per_step, fill=orc.fill_f32)
    for _ in range(args.warmup):
        m.forward(x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        m.forward(x)
    dt = time.perf_counter() - t0
    val = per_step * args.steps / dt
    sample = f"{per_step} of {args.batch} images per step (bounded CPU sample), CPU oracle port, OpenMP {threads} threads"
    print_json({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "s8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample,
                   "note": "the reference repo has no CPU or INT8 implementation of this path (SURVEY §0); "
                           "this arm is the CPU oracle port"},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


CONV_SLOTS = None      # indices of the 20 conv launches among the 23 profile slots (set from LAUNCH_NAMES)
BW_ALGO_BYTES = {      # algorithmic bytes per image of the bandwidth kernels
    "quantize_s2d": 602112 + 150528,      # fp32 image read + int8 image written (the s2d layout really writes 412160)
    "maxpool": 802816 + 200704,
    "gap_fc": 25088 + 4000,               # (+ 512 KB of FC weights per launch, amortised over the batch)
}


def _union_ms(intervals):
    """length of the union of [entry, exit] ns intervals, in ms"""
    iv = sorted((int(a), int(b)) for a, b in intervals if b > a)
    tot, cur_a, cur_b = 0, None, None
    for a, b in iv:
        if cur_b is None or a > cur_b:
            if cur_b is not None:
                tot += cur_b - cur_a
            cur_a, cur_b = a, b
        else:
            cur_b = max(cur_b, b)
    if cur_b is not None:
        tot += cur_b - cur_a
    return tot * 1e-6


def in_step_spans(stamps, names):
    """stamps: uint64 [F, L, 2] globaltimer ns of F consecutive forwards.  Returns per-forward dicts: the union of the
    conv launches' [first block entry, last block exit] intervals (the time conv kernels occupy the GPU inside the
    step, overlap with their neighbours counted once), each bandwidth kernel's span, and the forward's period."""
    import numpy as np
    conv_idx = [i for i, n in enumerate(names) if n not in BW_ALGO_BYTES]
    out = []
    F = stamps.shape[0]
    for f in range(F):
        st = stamps[f].astype(np.int64)
        valid = stamps[f, :, 0] != np.uint64(0xFFFFFFFFFFFFFFFF)
        conv_iv = [(st[i, 0], st[i, 1]) for i in conv_idx if valid[i]]
        d = {"conv_union_ms": _union_ms(conv_iv),
             "conv_sum_ms": float(sum(b - a for a, b in conv_iv)) * 1e-6,
             "forward_span_ms": float(st[valid, 1].max() - st[valid, 0].min()) * 1e-6}
        for i, n in enumerate(names):
            if n in BW_ALGO_BYTES and valid[i]:
                d[n + "_ms"] = float(st[i, 1] - st[i, 0]) * 1e-6
        if f + 1 < F:
            d["period_ms"] = float(int(stamps[f + 1, 0, 0]) - int(stamps[f, 0, 0])) * 1e-6
        t0 = st[valid, 0].min()
        d["spans_us"] = [[n, round(float(st[i, 0] - t0) * 1e-3, 2), round(float(st[i, 1] - t0) * 1e-3, 2)]
                         for i, n in enumerate(names) if valid[i]]
        out.append(d)
    return out


def bind_to_gpu_numa_node(index: int):
    """pin this process to the CPUs NVML lists for the GPU (pinned host buffers are then first-touched on the GPU's own
    NUMA node); returns the previous affinity so that the CPU baseline can restore it"""
    prev = None
    try:
        prev = os.sched_getaffinity(0)
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
        nv.nvmlDeviceSetCpuAffinity(nv.nvmlDeviceGetHandleByIndex(phys))
    except Exception:
        pass
    return prev


def reference_gpu_fp32(tmp_root: str):
    """SURVEY 8d(iii) / BASELINE.md 4 row 3: the reference's REAL implementation of this path - its FP32 step8_e2e binary
    (R/infer_e2e.cu:230-441, compiled unmodified into oracle/_ref/) - on this box's GPU: one image per process, wall clock
    per process, which is exactly what the reference's own harness measures (T/bench_fp32_vs_torch_e2e.py:105-120)."""
    import numpy as np
    import dlq_b200
    from dlq_b200 import synth
    exe = os.path.join(ROOT, "oracle", "_ref", "step8_e2e")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/step8_e2e not built"}
    d = os.path.join(tmp_root, "ref_w")
    dlq_b200.save_weight_dir(d, synth.make_weights(0))
    inp = os.path.join(tmp_root, "ref_input.bin")
    synth.make_input(0, 1).tofile(inp)
    walls = []
    top1 = None
    for _ in range(4):
        t0 = time.perf_counter()
        r = subprocess.run([exe, "--manifest", d, "--input", inp], capture_output=True, text=True, timeout=120)
        walls.append((time.perf_counter() - t0) * 1e3)
        if r.returncode != 0:
            return {"unavailable": f"step8_e2e exit {r.returncode}: {r.stderr[-200:]}"}
        for line in r.stdout.splitlines():
            if "top-1" in line:
                top1 = line.strip()
    out = {"impl": "oracle/_ref/step8_e2e = R/infer_e2e.cu + K/*.cu, unmodified, sm_100a", "batch": 1,
           "wall_ms_per_process": {"first": walls[0], "median_of_rest": float(np.median(walls[1:]))},
           "images_per_s": 1e3 / float(np.median(walls[1:])), "stdout": top1,
           "how": "one process per image (CUDA context + 45 MB of weight files + 88 launches), as the reference's bench script runs it"}
    prof = os.path.join(ROOT, "profiles", "r02_ref_step8_gpu.json")
    if os.path.exists(prof):
        with open(prof) as f:
            out["gpu_section_from_profile"] = json.load(f)
    return out


def run_ours(args, rank: int, local_rank: int, world: int):
    import numpy as np
    import torch
    import dlq_b200
    from dlq_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    prev_affinity = bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B = args.batch
    ctx = dlq_b200.Context(local_rank)
    weights = synth.make_weights(0)
    scales = synth.load_act_scales(0)
    model = dlq_b200.ResNet18(ctx, weights, scales, B)
    if args.chain_launch_mode:
        model.set_option("chain_launch_mode", args.chain_launch_mode)
    names = model.LAUNCH_NAMES
    # every rank gets its own shard of the global synthetic batch (weak scaling: B images per rank)
    xh_np = synth.make_input(rank, 8)
    xh = torch.from_numpy(np.ascontiguousarray(np.tile(xh_np, (B // 8 + 1, 1, 1, 1))[:B])).pin_memory()
    lh = torch.empty((B, 1000), dtype=torch.float32).pin_memory()
    lh2 = torch.empty((B, 1000), dtype=torch.float32).pin_memory()
    x = xh.cuda()
    logits = torch.empty((B, 1000), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    ctx.auto_order = False        # every buffer is complete from here on: no per-call event traffic in the timed loops
    stream = torch.cuda.ExternalStream(ctx.stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sampler=None, finish=None):
        for _ in range(warmup):
            fn()
        if finish:
            finish()
        ctx.sync()
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        if finish:
            finish()
        e1.record(stream)
        ctx.sync()
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, clocks

    def timed_wall(fn, steps, warmup, finish):
        """host-buffer pipelines end on other streams than the library's: wall clock between barriers"""
        for _ in range(warmup):
            fn()
        finish()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        finish()
        barrier()
        ms = (time.perf_counter() - t0) * 1e3
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident throughput (input batch 154 MB + ~1 GB of activations per step >> 126 MB L2)
    sampler = ClockSampler(local_rank)
    W = max(3, args.warmup)
    ms, clocks = timed(lambda: model.forward(x, logits), args.steps, W, sampler)
    value = B * world * args.steps / (ms * 1e-3)
    step_ms = ms / args.steps
    ref_logits = logits.clone()

    # ---- the same steps again with span stamps on: every kernel records the globaltimer of its first block entry
    # and last block exit, i.e. its span INSIDE the running step, programmatic-dependent-launch overlap included
    ring = min(args.steps, 64)
    model.enable_stamps(ring)
    ms_st, _ = timed(lambda: model.forward(x, logits), ring, 0)
    stamps = model.read_stamps()
    model.enable_stamps(0)
    spans = in_step_spans(stamps, names)[2:] if stamps.shape[0] > 4 else in_step_spans(stamps, names)
    med = lambda k: float(np.median([d[k] for d in spans if k in d]))
    conv_union_ms = med("conv_union_ms")
    # (forwards that are not back to back - a host synchronisation in between - have no meaningful period)
    periods = [d["period_ms"] for d in spans if "period_ms" in d and d["period_ms"] < 10 * step_ms]
    period_ms = float(np.median(periods)) if periods else step_ms
    rep = sorted(spans, key=lambda d: d["conv_union_ms"])[len(spans) // 2]       # the median forward, launch by launch

    # ---- per-launch durations with a CUDA event between launches (forbids the overlap the step runs with)
    prof = np.zeros(model.launches, dtype=np.float64)
    reps = 5
    model.profile(x, logits)
    for _ in range(reps):
        prof += model.profile(x, logits)
    prof /= reps
    conv_idx = [i for i, n in enumerate(names) if n not in BW_ALGO_BYTES]
    conv_ms_serial = float(sum(prof[i] for i in conv_idx))

    peaks, peak_kind = measured_peaks()
    sm_mhz = (clocks or {}).get("sm_mhz") or 0.0
    burst = sm_mhz >= 1900.0
    peak_tops = 2.0 * float(peaks["bf16_tflops"] if burst else peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    achieved = CONV_GOP_PER_IMG * B / (conv_union_ms * 1e-3) / 1e3      # TOP/s
    n_chain_layers, n_chains = model.plan_info(B, "chain_layers"), model.plan_info(B, "chains")
    conv_kernels_desc = (f"the 20 convs: conv_i8_kernel ({20 - n_chain_layers} launch(es)/step: the stem"
                         f"{'' if n_chain_layers >= 19 else ' and the convs outside the chains'}) + conv_chain_kernel "
                         f"({n_chains} launch(es)/step running {n_chain_layers} layers: layer1 as a single-CTA chain, layer2..4 as a CTA-pair chain)")
    # the tcgen05 kind::i8 issue rate this repo measured on a B200 (probe/mma_rate.cu): M=256 N=256 K=32 per 128.1 clk
    probe_mac_clk_sm = 256 * 256 * 32 / 128.1 / 2
    probe_tops = probe_mac_clk_sm * 2 * 148 * (sm_mhz or 1965.0) * 1e6 / 1e12

    # ---- end to end through the host-buffer entry points (pinned HOST fp32 batch, H2D + D2H inside the timed region)
    e2e_steps = max(3, min(args.steps, 20))

    def pipe(submit):
        state = {"n": 0}

        def step():
            submit()
            if state["n"] > 0:
                model.wait()
            state["n"] += 1

        def finish():
            model.wait()
            model.wait()
            state["n"] = 0
        return step, finish
    st, fin = pipe(lambda: model.submit_host(xh, lh))
    ms_e2e = timed_wall(st, e2e_steps, 3, fin)
    e2e_value = B * world * e2e_steps / (ms_e2e * 1e-3)
    e2e_ok = bool(torch.equal(lh.view(torch.int32), ref_logits.cpu().view(torch.int32)))
    ms_sync = timed_wall(lambda: model.forward_host(xh, lh2), e2e_steps, 2, lambda: None)
    e2e_sync = B * world * e2e_steps / (ms_sync * 1e-3)
    e2e_ok = e2e_ok and bool(torch.equal(lh2.view(torch.int32), ref_logits.cpu().view(torch.int32)))

    # ---- the same from uint8 images (device-side normalise + quantise table): 4x fewer H2D bytes
    e2e_u8 = None
    try:
        model.set_preprocess()
        rng_u8 = np.random.default_rng(rank)
        uh = torch.from_numpy(rng_u8.integers(0, 256, (B, 224, 224, 3), dtype=np.uint8)).pin_memory()
        st, fin = pipe(lambda: model.submit_host_u8(uh, lh))
        ms_u8_trials = sorted(timed_wall(st, e2e_steps, 3, fin) for _ in range(3))
        ms_u8 = ms_u8_trials[1]          # median of three trials: this number moves with the host side of the box
        du = uh.cuda()
        lu = torch.empty((B, 1000), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        model.forward_u8(du, lu)
        ctx.sync()
        e2e_u8 = {"value": B * world * e2e_steps / (ms_u8 * 1e-3), "unit": "images/s",
                  "h2d_bytes_per_step": int(uh.numel()) * world, "d2h_bytes_per_step": int(lh.numel() * 4) * world,
                  "logits_verified": bool(torch.equal(lh.view(torch.int32), lu.cpu().view(torch.int32))),
                  "trials_ms_per_step": [round(t / e2e_steps, 4) for t in ms_u8_trials],
                  "note": "pinned host uint8 HWC images -> H2D -> normalise+quantise+forward -> D2H logits; pipelined "
                          "(dlq_resnet18_submit_host_u8 / _wait: copy of batch k+1 under the forward of batch k); median of "
                          "three trials"}
        del du, lu
    except Exception as ex:
        e2e_u8 = {"value": None, "error": str(ex)}

    # ---- A/B on this box, in this process: the conv chain (default) vs one launch per conv with grid-level dependencies
    chain_ab = None
    if not args.no_extras:
        try:
            model.set_option("conv_chain", 0)
            ms_grid, _ = timed(lambda: model.forward(x, logits), args.steps, 3)
            same = bool(torch.equal(logits.view(torch.int32), ref_logits.view(torch.int32)))
            launches_grid = model.launches_for_batch(B)
            model.set_option("conv_chain", 1)
            ms_chain, _ = timed(lambda: model.forward(x, logits), args.steps, 3)
            chain_ab = {"ms_per_step_chain": ms_chain / args.steps, "ms_per_step_one_launch_per_conv": ms_grid / args.steps,
                        "launches_chain": model.launches_for_batch(B), "launches_one_per_conv": launches_grid,
                        "chain_layers": model.plan_info(B, "chain_layers"), "logits_identical": same,
                        "dep_timeouts": model.dep_timeouts,
                        "how": "dlq_resnet18_set_option(conv_chain): the thirteen convs from layer2.0.conv2 on as ONE persistent "
                               "cooperative kernel (csrc/conv_chain.cuh) vs one launch each, same process, same box"}
        except Exception as ex:
            chain_ab = {"error": str(ex)}

    # ---- sustained: the same step for >= 3 s (clocks settle to the sustained level), against the sustained peak
    sustained = None
    if not args.no_extras:
        try:
            n_s = max(50, int(3000.0 / step_ms))
            samp2 = ClockSampler(local_rank)
            ms_s, clk_s = timed(lambda: model.forward(x, logits), n_s, 0, samp2)
            conv_share = conv_union_ms / period_ms
            conv_ms_s = ms_s / n_s * conv_share
            pk_s = 2.0 * float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
            sustained = {"seconds": ms_s * 1e-3, "steps": n_s, "value": B * world * n_s / (ms_s * 1e-3), "unit": "images/s",
                         "ms_per_step": ms_s / n_s, "clocks": clk_s, "conv_share_of_step": conv_share,
                         "conv_achieved_tops": CONV_GOP_PER_IMG * B / (conv_ms_s * 1e-3) / 1e3, "peak_tops": pk_s,
                         "conv_frac": CONV_GOP_PER_IMG * B / (conv_ms_s * 1e-3) / 1e3 / pk_s,
                         "how": "conv time = step time x the conv family's share of the step from the span stamps"}
        except Exception as ex:
            sustained = {"error": str(ex)}

    # ---- BASELINE config 4: the E4M3 network on the same batch (device-resident), config 2: batch-1 latency
    fp8_info, lat_info = None, None
    if not args.no_extras:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            s_int8 = np.asarray(scales, dtype=np.float64)
            s_fp8 = (s_int8 * 127.0 / 448.0).astype(np.float32)     # same calibrated absmax, mapped to 448
            m8 = dlq_b200.ResNet18(ctx, weights, s_fp8, B, fp8=True)
            n8 = max(5, args.steps // 2)
            ms8, _ = timed(lambda: m8.forward(x, logits), n8, 3)
            fp8_info = {"value": B * world * n8 / (ms8 * 1e-3), "unit": "images/s",
                        "dtype": "e4m3 x e4m3 -> f32 (tcgen05 kind::f8f6f4)", "batch_per_gpu": B,
                        "vs_int8": (B * world * n8 / (ms8 * 1e-3)) / value,
                        "why_slower": "same kernels and tile shapes (tools/net_spans.py): layer1-2 equal, every layer3 / layer4 "
                                      "conv 4-6 us longer (the tensor-bound layers; a tight-loop probe issues kind::f8f6f4 and kind::i8 pair MMAs at the same "
                                      "rate, so the cause is not the MMA rate itself and was not isolated), and "
                                      "the GAP+FC tail 16 us longer (FP32 sums in the oracle's sequential order, E4M3 weights "
                                      "decoded per use, where the int8 tail is dp4a)"}
            m8.close()
        except Exception as ex:
            fp8_info = {"value": None, "error": str(ex)}
        if world == 1:
            try:
                m1 = dlq_b200.ResNet18(ctx, weights, scales, 1)
                x1, l1 = x[:1].contiguous(), torch.empty((1, 1000), dtype=torch.float32, device="cuda")
                torch.cuda.synchronize()
                m1.graph_capture(x1, l1)
                for _ in range(20):
                    m1.graph_launch()
                ctx.sync()
                evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(300)]
                for a, b in evs:
                    a.record(stream)
                    m1.graph_launch()
                    b.record(stream)
                ctx.sync()
                t = np.array([a.elapsed_time(b) for a, b in evs]) * 1e3
                lat_info = {"batch": 1, "median_us": float(np.median(t)), "p99_us": float(np.percentile(t, 99)),
                            "launches": m1.launches_for_batch(1),
                            "how": "CUDA-graph replay of one forward, device-resident fp32 input, 300 replays"}
                m1.close()
            except Exception as ex:
                lat_info = {"error": str(ex)}
    # ---- BASELINE config 4's per-GPU batch at two GPUs (2048 / 2): the same measurement at batch 1024
    big_info = None
    if not args.no_extras and world == 1 and B == 256:
        try:
            Bb = 1024
            mb = dlq_b200.ResNet18(ctx, weights, scales, Bb)
            xb = x.repeat(4, 1, 1, 1).contiguous()
            lb = torch.empty((Bb, 1000), dtype=torch.float32, device="cuda")
            torch.cuda.synchronize()
            nb = max(5, args.steps // 4)
            msb, _ = timed(lambda: mb.forward(xb, lb), nb, 3)
            mb.enable_stamps(8)
            timed(lambda: mb.forward(xb, lb), 8, 0)
            sb = in_step_spans(mb.read_stamps(), names)[2:]
            mb.enable_stamps(0)
            conv_b = float(np.median([d["conv_union_ms"] for d in sb]))
            ach_b = CONV_GOP_PER_IMG * Bb / (conv_b * 1e-3) / 1e3
            big_info = {"batch": Bb, "value": Bb * nb / (msb * 1e-3), "unit": "images/s", "ms_per_step": msb / nb,
                        "conv_ms_in_step": conv_b, "conv_achieved_tops": ach_b, "conv_frac_of_roofline": ach_b / peak_tops,
                        "how": "device-resident, same method as the headline line (union of the conv launches' in-step spans), "
                               "batch 1024 = BASELINE config 4's per-GPU share at two GPUs"}
            mb.close()
            del xb, lb
            torch.cuda.empty_cache()
        except Exception as ex:
            big_info = {"error": str(ex)}

    # ---- accuracy harness on synthetic images: INT8 / FP8 logits vs the reference's FP32 arithmetic
    # (dlq_resnet18_f32_*, bit-exact restatement of the reference's operators), in-process
    acc_info = None
    if not args.no_extras and world == 1:
        try:
            ctx.auto_order = True
            na = 32
            xa = torch.from_numpy(synth.make_input(21, na)).cuda()
            fnet = dlq_b200.ResNet18F32(ctx, weights, na)
            lf = torch.empty((na, 1000), dtype=torch.float32, device="cuda")
            lq = torch.empty((na, 1000), dtype=torch.float32, device="cuda")
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            fnet.forward(xa, lf)
            ev0.record(stream)
            fnet.forward(xa, lf)
            ev1.record(stream)
            ctx.sync()
            tf = torch.empty((na, 1), dtype=torch.int32, device="cuda")
            tq = torch.empty((na, 1), dtype=torch.int32, device="cuda")
            ctx.topk_f32(lf, 1, tf)
            acc_info = {"images": na, "fp32_reference_arith_images_per_s": na / (ev0.elapsed_time(ev1) * 1e-3),
                        "how": "logits of the quantised network vs dlq_resnet18_f32_forward on the same synthetic images"}
            for tag, is8 in (("int8", False), ("fp8", True)):
                sc = np.asarray(scales, dtype=np.float64)
                sc = (sc * 127.0 / 448.0).astype(np.float32) if is8 else sc.astype(np.float32)
                mq = dlq_b200.ResNet18(ctx, weights, sc, na, fp8=is8)
                mq.forward(xa, lq)
                r = ctx.compare_f32(lf, lq)
                ctx.topk_f32(lq, 1, tq)
                ctx.sync()
                acc_info[tag] = {"cosine": r["cosine"], "max_abs": r["max_abs"], "mean_abs": r["mean_abs"],
                                 "agree_top1": float((tf == tq).float().mean().item())}
                mq.close()
            fnet.close()
        except Exception as ex:
            acc_info = {"error": str(ex)}

    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            if prev_affinity:
                os.sched_setaffinity(0, prev_affinity)       # the CPU baseline may use every host core
            rate, threads, dt = cpu_oracle_rate(args.cpu_sample)
            cpu = {"value": rate, "unit": "images/s", "cores": threads, "kind": "port",
                   "sample": f"{args.cpu_sample} of {B} images, CPU oracle INT8 forward (restatement + QUANT_SPEC), "
                             f"{dt:.1f} s; the reference has no CPU ResNet path"}
        except Exception as ex:   # the oracle is a checker; never let it break the product measurement
            cpu = {"value": None, "unit": "images/s", "cores": 0, "kind": "port", "sample": f"unavailable: {ex}"}

    mnist_info, ref_gpu = None, None
    if world == 1 and not args.no_extras:
        try:
            mnist_info = mnist_config0(ctx)
        except Exception as ex:
            mnist_info = {"error": str(ex)}
        try:
            import tempfile
            with tempfile.TemporaryDirectory() as td:
                ref_gpu = reference_gpu_fp32(td)
        except Exception as ex:
            ref_gpu = {"unavailable": str(ex)}

    roofline_bw = {}
    for i, n in enumerate(names):
        if n in BW_ALGO_BYTES:
            t_in = med(n + "_ms") if any((n + "_ms") in d for d in spans) else None
            gbs = BW_ALGO_BYTES[n] * B / (t_in * 1e-3) / 1e9 if t_in else None
            roofline_bw[n] = {"bound": "hbm", "algorithmic_bytes_per_launch": BW_ALGO_BYTES[n] * B,
                              "ms_in_step": t_in, "ms_between_events": float(prof[i]), "achieved": gbs,
                              "peak": float(peaks["hbm_gbs"]), "unit": "GB/s", "frac": gbs / float(peaks["hbm_gbs"]) if gbs else None,
                              "traffic": (traffic or {}).get("dram_bytes_per_launch", {}).get(n)}
    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": W, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "s8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B * world, "sharding": f"batch x{world}",
                   "l2": "inputs larger than L2 (154 MB fp32 batch + ~1 GB activations per step vs 126 MB L2)"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": int(xh.numel() * 4) * world,
                "d2h_bytes_per_step": int(lh.numel() * 4) * world, "steps": e2e_steps, "logits_verified": e2e_ok,
                "synchronous_value": e2e_sync,
                "note": "pinned host fp32 batch (the reference's input format) -> H2D -> forward -> D2H logits, every step, "
                        "through dlq_resnet18_submit_host / _wait (copy of batch k+1 under the forward of batch k; wall clock "
                        "between barriers); 'synchronous_value' = one dlq_resnet18_forward_host call per step. PCIe-bound: "
                        "602 KB per image"},
        "gpu_launches": int(model.launches_for_batch(B) * args.steps),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tops, "unit": "TFLOP/s", "frac": achieved / peak_tops,
                     "traffic": (traffic or {}).get("conv_family_dram_bytes_per_step"),
                     "traffic_source": (traffic or {}).get("source"),
                     "kernel": conv_kernels_desc,
                     "peak_source": f"2 x bf16_tflops{'' if burst else '_sustained'} ({peak_kind}); int8 dense = 2 x bf16 on sm_100; "
                                    f"burst figure because the sampled SM clock is {sm_mhz:.0f} MHz" if burst else
                                    f"2 x bf16_tflops_sustained ({peak_kind}): sampled SM clock {sm_mhz:.0f} MHz < 1900",
                     "ops": "int8 MAC*2", "algorithmic_gop_per_step": CONV_GOP_PER_IMG * B,
                     "conv_ms_in_step": conv_union_ms,
                     "conv_ms_method": "union of the 20 conv launches' [first block entry, last block exit] globaltimer spans "
                                       "inside the running step (stamps; median over the stamped forwards) - overlap of "
                                       "neighbouring launches counted once, waiting at griddepcontrol.wait included",
                     "conv_ms_serialised": conv_ms_serial, "step_period_ms_stamped": period_ms,
                     "frac_if_convs_charged_the_whole_step": CONV_GOP_PER_IMG * B / (step_ms * 1e-3) / 1e3 / peak_tops,
                     "frac_vs_serialised": CONV_GOP_PER_IMG * B / (conv_ms_serial * 1e-3) / 1e3 / peak_tops,
                     "peak_i8_probe": {"mac_per_clk_per_sm": probe_mac_clk_sm, "tops_at_sampled_clock": probe_tops,
                                       "frac": achieved / probe_tops,
                                       "source": "probe/mma_rate.cu on a B200 (profiles/r01_mma_rate.log): cta_group::2 "
                                                 "M=256 N=256 K=32 kind::i8 issues every 128.1 clk"},
                     "frac_of_vendor_dense_spec": {"peak_tops": 4500.0, "frac": achieved / 4500.0,
                                                   "note": "NVIDIA's dense INT8 figure for B200 (SURVEY 8d asks for both denominators)"},
                     "hbm_gbs_conv": CONV_BYTES_PER_IMG * B / (conv_union_ms * 1e-3) / 1e9,
                     "in_step_spans_us": rep["spans_us"],
                     "per_launch_ms_between_events": {n: round(float(v), 4) for n, v in zip(names, prof)}},
        "roofline_bw": roofline_bw,
        "chain_ab": chain_ab,
        "sustained": sustained,
        "cpu_baseline": cpu,
        "e2e_u8": e2e_u8,
        "mnist_config0": mnist_info,
        "reference_gpu_fp32": ref_gpu,
        "fp8": fp8_info,
        "latency_b1": lat_info,
        "accuracy_vs_fp32": acc_info,
        "batch_1024": big_info,
    }
    print_json(line)
    if dist is not None:
        dist.destroy_process_group()


def _claim_stdout():
    """Route everything libraries print on fd 1 (e.g. the NCCL version banner) to stderr and return a writer
    for the real stdout, so that stdout carries exactly one JSON line."""
    real = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    return real


def main():
    real_stdout = _claim_stdout()
    global print_json

    def print_json(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--cpu-sample", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the FP8 and batch-1 latency side measurements")
    ap.add_argument("--chain-launch-mode", type=int, default=0, choices=[0, 1, 2],
                    help="launch attributes of the conv chain kernel: 0 cooperative + programmatic serialization (default), "
                         "1 cooperative, 2 neither - for captures under ncu, which fails on cooperative cluster launches")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
