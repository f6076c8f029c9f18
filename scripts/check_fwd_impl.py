"""Forward of one level shape with a given fwd_impl against the oracle, plus its time.  usage: check_fwd_impl.py IMPL B H W C dtype [flow] (env R = search range, default 4)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pwc_oracle as O
from tfoptflow_b200 import _ffi, synth
from tfoptflow_b200.ops import warp_cost_volume_forward
impl, B, H, W, C = [int(x) for x in sys.argv[1:6]]
dt = {"fp32": np.float32, "fp16": np.float16}[sys.argv[6]]
fk = sys.argv[7] if len(sys.argv) > 7 else "smooth"
fk = None if fk == "none" else fk
R = int(os.environ.get("R", 4))
dev = torch.device("cuda", 0)
d = synth.make_level(B, H, W, C, 11, dt, fk, search_range=R)
t = {k: (None if d[k] is None else torch.from_numpy(d[k]).to(dev)) for k in ("c1", "c2", "flow")}
nchk = min(B, 2)
ref = O.warp_cost_volume(d["c1"][:nchk], d["c2"][:nchk], None if fk is None else d["flow"][:nchk], R, 2.5).astype(np.float64)
for im in (impl, 0):
    _ffi.set_option("fwd_impl", im)
    out = warp_cost_volume_forward(t["c1"], t["c2"], t["flow"], R, 2.5)
    torch.cuda.synchronize()
    got = out[:nchk].float().cpu().numpy().astype(np.float64)
    rms = np.sqrt(np.mean(ref * ref))
    e = np.abs(got - ref)
    for _ in range(3):
        warp_cost_volume_forward(t["c1"], t["c2"], t["flow"], R, 2.5, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        warp_cost_volume_forward(t["c1"], t["c2"], t["flow"], R, 2.5, out=out)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    byt = B * H * W * (2 * C + (2 if fk else 0) + 81) * d["c1"].itemsize
    print(f"fwd_impl {im} {B}x{H}x{W}x{C} {sys.argv[6]} {fk}: max err {e.max():.3e} = {e.max() / rms:.2e} x rms (rel max {np.max(e / (np.abs(ref) + rms)):.2e}); "
          f"{us:.1f} us, {byt / us / 1e3:.0f} GB/s algorithmic = {byt / us / 1e3 / 6553:.3f}", flush=True)
