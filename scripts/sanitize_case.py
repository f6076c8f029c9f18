"""Small streaming-kernel cases for compute-sanitizer (memcheck): fp32 + fp16, ragged width, strided output."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tfoptflow_b200 import _ffi, synth, warp_cost_volume
dev = torch.device("cuda", 0)
_ffi.set_option("fwd_impl", 3)
for (b, h, w, c, dt) in ((2, 44, 70, 32, np.float32), (1, 30, 37, 64, np.float32), (1, 44, 50, 32, np.float16), (1, 28, 20, 96, np.float32)):
    d = synth.make_level(b, h, w, c, 11, dt, "smooth")
    t = lambda a: torch.from_numpy(a).to(dev)
    o = warp_cost_volume(t(d["c1"]), t(d["c2"]), t(d["flow"]), 4, 2.5)
    torch.cuda.synchronize()
    print(b, h, w, c, np.dtype(dt).name, float(o.float().abs().max()), flush=True)
print("done")
