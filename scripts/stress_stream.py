"""Protocol stress for the streaming kernel: odd CTA counts, ragged shapes, stress flows, repeated launches; every
result is compared with the generic kernel (bit-level agreement up to the documented tolerance).  A hang shows up as
the caller's timeout."""
import os, sys, itertools
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tfoptflow_b200 import _ffi, synth, warp_cost_volume
dev = torch.device("cuda", 0)
shapes = [(3, 60, 50, 32), (2, 57, 129, 64), (5, 30, 24, 96), (2, 41, 38, 128), (7, 113, 25, 16), (1, 224, 512, 32)]
bad = 0
for (b, h, w, c), dt, fk in itertools.product(shapes, (np.float32, np.float16), ("stress", "smooth")):
    d = synth.make_level(b, h, w, c, 3 + h, dt, fk)
    t = lambda a: torch.from_numpy(a).to(dev)
    c1, c2, fl = t(d["c1"]), t(d["c2"]), t(d["flow"])
    _ffi.set_option("fwd_impl", 1); _ffi.set_option("stream_ctas", 0)
    ref = warp_cost_volume(c1, c2, fl, 4, 1.7).float()
    _ffi.set_option("fwd_impl", 3)
    for ctas in (0, 1, 5, 37, 111):
        print("case", (b, h, w, c), np.dtype(dt).name, fk, ctas, flush=True)
        _ffi.set_option("stream_ctas", ctas)
        for rep in range(2):
            got = warp_cost_volume(c1, c2, fl, 4, 1.7).float()
            torch.cuda.synchronize()
            err = float((got - ref).abs().max()); tol = (2e-6 if dt == np.float32 else 4e-3) * float(ref.abs().max())
            if not err <= tol:
                bad += 1
                print("BAD", (b, h, w, c), np.dtype(dt).name, fk, ctas, err, tol, flush=True)
print("stress done, bad =", bad, flush=True)
