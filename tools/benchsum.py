import json,sys
d=json.loads(sys.stdin.read())
r=d["roofline"]
print("img/s %.0f  ms/step %.4f  conv_ms %.4f  frac %.3f  e2e %.0f" % (d["value"], d["ms_per_step"], r["conv_ms_per_step"], r["frac"], d["e2e"]["value"]))
pl=r["per_launch_ms"]
print(" ".join(f"{k.replace('layer','L').replace('downsample','ds').replace('conv','c')}={v*1000:.0f}" for k,v in pl.items()))
