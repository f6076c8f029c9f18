"""Times the backward of every level of the headline workload (BASELINE configs[2] shapes, batch B) through the C ABI
with pre-allocated buffers (CUDA events, L2 flushed between launches by rotating over > 126 MB of tensors is not
needed: one level's tensors are 0.3 - 1.4 GB).  usage: time_bwd_levels.py [B] [dtype] [key=value ...]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tfoptflow_b200 import _ffi, synth
from tfoptflow_b200._tensors import dtype_code, ptr, stream_ptr
from tfoptflow_b200.ops import warp_cost_volume_forward

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dt = {"fp32": torch.float32, "fp16": torch.float16}[sys.argv[2] if len(sys.argv) > 2 else "fp32"]
opts = [kv.split("=") for kv in sys.argv[3:]]
lib = _ffi.load()
dev = torch.device("cuda", 0)
peak = 6553.0
tot = {}
for impl in ((0,) if os.environ.get("PWC_BWD_ONLY_FUSED") else (1, 0)):
    _ffi.set_option("bwd_impl", impl)
    for k, v in opts:
        _ffi.set_option(k, int(v))
    tot[impl] = 0.0
    only = [int(x) for x in os.environ.get("PWC_LEVELS", "6,5,4,3,2").split(",")]
    for (lvl, h, w, c) in synth.level_shapes(448, 1024, tuple(sorted(only, reverse=True))):
        g = torch.Generator(device=dev).manual_seed(lvl)
        c1 = torch.randn((B, h, w, c), device=dev, generator=g).to(dt)
        c2 = torch.randn((B, h, w, c), device=dev, generator=g).to(dt)
        flow = None if lvl == 6 else (torch.randn((B, h, w, 2), device=dev, generator=g) * 0.5).to(dt)
        gr = torch.randn((B, h, w, 81), device=dev, generator=g).to(dt)
        fs = synth.FLOW_SCALERS[lvl] if flow is not None else 1.0
        out = warp_cost_volume_forward(c1, c2, flow, 4, fs)
        dc1, dc2 = torch.empty_like(c1), torch.empty_like(c2)
        dfl = torch.empty_like(flow) if flow is not None else None
        nb = lib.pwc_bwd_workspace_bytes(B, h, w, c, 4, dtype_code(c1), 1 if flow is not None else 0)
        ws = torch.empty(max(nb, 256), dtype=torch.uint8, device=dev)
        def run():
            _ffi.check(lib.pwc_warp_cost_volume_bwd(ptr(gr), ptr(out), ptr(c1), ptr(c2), ptr(flow), fs, ptr(dc1), ptr(dc2),
                                                    ptr(dfl), B, h, w, c, 4, dtype_code(c1), dtype_code(c1), 0, 0, ptr(ws), nb,
                                                    stream_ptr(dev)))
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n): run()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        es = c1.element_size()
        byt = B * h * w * (2 * 81 + 4 * c + (4 if flow is not None else 0)) * es
        tot[impl] += us
        print(f"bwd_impl {impl} level {lvl} {B}x{h}x{w}x{c} {sys.argv[2] if len(sys.argv) > 2 else 'fp32'}: {us:8.1f} us  "
              f"{byt / us / 1e3:7.1f} GB/s algorithmic = {byt / us / 1e3 / peak:.3f} of {peak:.0f}", flush=True)
    print(f"bwd_impl {impl} all levels: {tot[impl]:.1f} us")
