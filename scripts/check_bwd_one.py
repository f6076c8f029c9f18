"""One fused-backward case against the oracle (quick sanity before the test suite).  usage: check_bwd_one.py B H W C [flow]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pwc_oracle as O
from tfoptflow_b200 import _ffi, synth, warp_cost_volume
B, H, W, C = [int(x) for x in sys.argv[1:5]]
fk = sys.argv[5] if len(sys.argv) > 5 else "smooth"
fk = None if fk == "none" else fk
dev = torch.device("cuda", 0)
d = synth.make_level(B, H, W, C, 7, np.float32, fk, with_grad=True)
ref = O.warp_cost_volume_bwd(d["g"], d["c1"], d["c2"], d["flow"], 4, 2.5)
for impl in (1, 2):
    _ffi.set_option("bwd_impl", impl)
    c1, c2 = (torch.from_numpy(d[k]).to(dev).requires_grad_() for k in ("c1", "c2"))
    fl = None if fk is None else torch.from_numpy(d["flow"]).to(dev).requires_grad_()
    warp_cost_volume(c1, c2, fl, 4, 2.5).backward(torch.from_numpy(d["g"]).to(dev))
    torch.cuda.synchronize()
    for name, got, want in (("dc1", c1.grad, ref[0]), ("dc2", c2.grad, ref[1]), ("dflow", None if fl is None else fl.grad, ref[2])):
        if got is None:
            continue
        e = np.abs(got.cpu().numpy() - want)
        rms = np.sqrt(np.mean(want.astype(np.float64) ** 2))
        print(f"impl {impl} {name}: max err {e.max():.3e} ({e.max() / rms:.2e} x rms) at {np.unravel_index(e.argmax(), e.shape)}", flush=True)
        if not e.max() <= 2e-5 * rms:
            raise SystemExit(f"impl {impl} {name}: error above 2e-5 x rms")
