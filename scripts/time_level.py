"""Times one pyramid level's fused forward with CUDA events.  usage: time_level.py [B] [impl] [key=value ...]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tfoptflow_b200 import _ffi, synth, pyramid
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
impl = int(sys.argv[2]) if len(sys.argv) > 2 else 3
for kv in sys.argv[3:]:
    k, v = kv.split("="); _ffi.set_option(k, int(v))
_ffi.set_option("fwd_impl", impl)
dev = torch.device("cuda", 0)
d = synth.make_level(B, 112, 256, 32, 1971, np.float32, "smooth")
L = {k: (None if d[k] is None else torch.from_numpy(d[k]).to(dev)) for k in ("c1", "c2", "flow")}
L["scaler"] = 5.0
plan = pyramid.PyramidPlan([L], 4)
for _ in range(5): plan.run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 30
e0.record()
for _ in range(n): plan.run()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / n * 1e3
byts = B * 112 * 256 * (64 + 2 + 81) * 4
print(f"B={B} impl={impl} opts={sys.argv[3:]}: {us:8.1f} us/launch  {us / B:6.2f} us/pair  {byts / us / 1e3:7.1f} GB/s algorithmic", flush=True)
