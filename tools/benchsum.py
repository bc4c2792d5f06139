#!/usr/bin/env python
"""one-screen summary of a bench.py line:  python tools/benchsum.py gpurun_out/x.json"""
import json
import sys

d = json.load(open(sys.argv[1]))
r = d["roofline"]
print("img/s %.0f  ms/step %.4f  conv_ms_in_step %.4f  frac %.3f (whole step %.3f, serialised %.3f)  clocks %s" % (
    d["value"], d["ms_per_step"], r["conv_ms_in_step"], r["frac"], r["frac_if_convs_charged_the_whole_step"], r["frac_vs_serialised"],
    d.get("clocks")))
print("e2e %.0f (sync %.0f, verified %s)  e2e_u8 %s" % (d["e2e"]["value"], d["e2e"]["synchronous_value"], d["e2e"]["logits_verified"],
                                                      (d.get("e2e_u8") or {}).get("value")))
for k in ("chain_ab", "sustained", "fp8", "latency_b1", "batch_1024", "cpu_baseline", "reference_gpu_fp32", "mnist_config0", "accuracy_vs_fp32"):
    print(k + ":", json.dumps(d.get(k))[:420])
print("bw:", {k: (round(v["frac"], 3) if v.get("frac") else None) for k, v in d.get("roofline_bw", {}).items()})
print("spans:", "  ".join("%s:%.0f-%.0f" % (n.replace("layer", "L"), a, b) for n, a, b in r["in_step_spans_us"]))
