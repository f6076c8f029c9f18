"""Times the fused forward for one level shape with CUDA events.  usage: time_shape.py B H W C impl [scale] [key=value ...]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tfoptflow_b200 import _ffi, synth, pyramid
B, H, W, C, impl = [int(x) for x in sys.argv[1:6]]
rest = sys.argv[6:]
has_flow = True
for kv in rest:
    k, v = kv.split("=")
    if k == "flow": has_flow = bool(int(v))
    else: _ffi.set_option(k, int(v))
_ffi.set_option("fwd_impl", impl)
dev = torch.device("cuda", 0)
d = synth.make_level(B, H, W, C, 1971, np.float32, "smooth" if has_flow else None)
L = {k: (None if d[k] is None else torch.from_numpy(d[k]).to(dev)) for k in ("c1", "c2", "flow")}
L["scaler"] = 1.25
plan = pyramid.PyramidPlan([L], 4)
for _ in range(5): plan.run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 30
e0.record()
for _ in range(n): plan.run()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / n * 1e3
byts = B * H * W * (2 * C + (2 if has_flow else 0) + 81) * 4
print(f"{B}x{H}x{W}x{C} impl={impl} flow={has_flow} {rest}: {us:8.1f} us/launch  {byts / us / 1e3:7.1f} GB/s algorithmic", flush=True)
