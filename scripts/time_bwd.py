"""Times forward + backward of the fused operator at one level shape (CUDA events).  usage: time_bwd.py B H W C"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tfoptflow_b200 import synth, warp_cost_volume
B, H, W, C = [int(x) for x in sys.argv[1:5]]
from tfoptflow_b200 import _ffi
for kv in sys.argv[5:]:
    k, v = kv.split('='); _ffi.set_option(k, int(v))
dev = torch.device("cuda", 0)
d = synth.make_level(B, H, W, C, 1971, np.float32, "smooth", with_grad=True)
c1, c2, flow = (torch.from_numpy(d[k]).to(dev).requires_grad_() for k in ("c1", "c2", "flow"))
g = torch.from_numpy(d["g"]).to(dev)
def step():
    for t in (c1, c2, flow): t.grad = None
    out = warp_cost_volume(c1, c2, flow, 4, 5.0)
    out.backward(g)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
fwd_b = B * H * W * (2 * C + 2 + 81) * 4; bwd_b = B * H * W * (2 * 81 + 4 * C + 4) * 4
print(f"{B}x{H}x{W}x{C}: fwd+bwd {ms*1e3:.1f} us  ({(fwd_b + bwd_b) / ms / 1e6:.1f} GB/s algorithmic over fwd+bwd bytes)")
