"""Times the backward of one shape for bwd_impl 1 (unfused) and 2 (fused).  usage: time_bwd_shape.py B H W C [noflow]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tfoptflow_b200 import _ffi
from tfoptflow_b200._tensors import dtype_code, ptr, stream_ptr
from tfoptflow_b200.ops import warp_cost_volume_forward
B, h, w, c = [int(x) for x in sys.argv[1:5]]
noflow = len(sys.argv) > 5 and sys.argv[5] == "noflow"
lib = _ffi.load(); dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
c1 = torch.randn((B, h, w, c), device=dev, generator=g); c2 = torch.randn((B, h, w, c), device=dev, generator=g)
flow = None if noflow else torch.randn((B, h, w, 2), device=dev, generator=g) * 0.5
gr = torch.randn((B, h, w, 81), device=dev, generator=g)
out = warp_cost_volume_forward(c1, c2, flow, 4, 2.5)
dc1, dc2 = torch.empty_like(c1), torch.empty_like(c2)
dfl = torch.empty_like(flow) if flow is not None else None
nb = lib.pwc_bwd_workspace_bytes(B, h, w, c, 4, dtype_code(c1), 0 if noflow else 1)
ws = torch.empty(max(nb, 256), dtype=torch.uint8, device=dev)
for impl in (1, 2):
    _ffi.set_option("bwd_impl", impl)
    def run():
        _ffi.check(lib.pwc_warp_cost_volume_bwd(ptr(gr), ptr(out), ptr(c1), ptr(c2), ptr(flow), 2.5, ptr(dc1), ptr(dc2), ptr(dfl),
                                                B, h, w, c, 4, dtype_code(c1), dtype_code(c1), 0, 0, ptr(ws), nb, stream_ptr(dev)))
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    print(f"{B}x{h}x{w}x{c} {'noflow' if noflow else 'flow'} bwd_impl {impl}: {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us", flush=True)
