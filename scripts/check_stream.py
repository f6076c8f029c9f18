"""Quick GPU check of the streaming kernel (fwd_impl=3) against the generic kernel (fwd_impl=1)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tfoptflow_b200 import _ffi, synth, warp_cost_volume
dev = torch.device("cuda", 0)
def run(d, impl, fs):
    _ffi.set_option("fwd_impl", impl)
    t = lambda a: None if a is None else torch.from_numpy(a).to(dev)
    o = warp_cost_volume(t(d["c1"]), t(d["c2"]), t(d["flow"]), 4, fs)
    torch.cuda.synchronize()
    return o.float().cpu().numpy()
cases = [(1, 8, 32, 32, "smooth"), (1, 12, 40, 32, "stress"), (2, 7, 33, 32, "smooth"), (3, 24, 64, 32, None),
         (2, 96, 128, 32, "smooth"), (1, 13, 70, 16, "stress"), (4, 112, 256, 32, "smooth"), (1, 2, 2, 32, "smooth"),
         (2, 56, 128, 64, "smooth"), (1, 9, 21, 64, "stress"), (3, 28, 64, 96, "smooth"), (2, 11, 9, 96, None),
         (4, 14, 32, 128, "smooth"), (1, 5, 19, 128, "stress"), (8, 56, 128, 64, None)]
for ctas in (0, 3):
    _ffi.set_option("stream_ctas", ctas)
    for (b, h, w, c, fk) in cases:
        for dt in (np.float32, np.float16):
            d = synth.make_level(b, h, w, c, 7 + h + w, dt, fk)
            t0 = time.time()
            ref = run(d, 1, 1.25)
            got = run(d, 3, 1.25)
            err = float(np.abs(got - ref).max())
            tol = (2e-6 if dt == np.float32 else 4e-3) * float(np.abs(ref).max())
            print(f"ctas={ctas} {b}x{h}x{w}x{c} {fk} {np.dtype(dt).name}: {'OK ' if err <= tol else 'BAD'} max|stream-generic| = {err:.3e}  "
                  f"nonzero={float(np.abs(ref).max()):.3f}  ({time.time() - t0:.2f}s)", flush=True)
