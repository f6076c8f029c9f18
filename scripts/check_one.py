import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tfoptflow_b200 import _ffi, synth, warp_cost_volume
dev = torch.device("cuda", 0)
b, h, w, c = [int(x) for x in sys.argv[1:5]]
d = synth.make_level(b, h, w, c, 7 + h + w, np.float32, "smooth")
t = lambda a: None if a is None else torch.from_numpy(a).to(dev)
def run(impl):
    _ffi.set_option("fwd_impl", impl)
    o = warp_cost_volume(t(d["c1"]), t(d["c2"]), t(d["flow"]), 4, 1.25); torch.cuda.synchronize(); return o.cpu().numpy()
ref = run(1)
for ctas in [int(x) for x in sys.argv[5:]]:
    _ffi.set_option("stream_ctas", ctas)
    errs = []
    for rep in range(3):
        got = run(3)
        e = np.abs(got - ref)
        bad = np.argwhere(e > 1e-5)
        errs.append((float(e.max()), len(bad), tuple(bad[0]) if len(bad) else None, tuple(bad[-1]) if len(bad) else None))
    print(f"{b}x{h}x{w}x{c} ctas={ctas}:", errs, flush=True)
