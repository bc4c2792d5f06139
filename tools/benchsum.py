import json, sys
d = json.loads(open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read())
r = d["roofline"]
print("img/s %.0f  ms/step %.4f  conv_ms %.4f  frac %.3f  e2e %.0f  clocks %s" % (d["value"], d["ms_per_step"], r["conv_ms_per_step"], r["frac"], d["e2e"]["value"], d.get("clocks")))
pl = r["per_launch_ms"]
print(" ".join(f"{k.replace('layer','L').replace('downsample','ds').replace('conv','c')}={v*1000:.0f}" for k, v in pl.items()))
print("fp8:", d.get("fp8"), "| latency_b1:", d.get("latency_b1"))
print("e2e_u8:", (d.get("e2e_u8") or {}).get("value"))
print("mnist_config0:", d.get("mnist_config0"))
print("accuracy_vs_fp32:", d.get("accuracy_vs_fp32"))
print("batch_1024:", d.get("batch_1024"))
