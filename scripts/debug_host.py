import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tfoptflow_b200 import pyramid, synth, warp_cost_volume, _ffi
dev = torch.device("cuda", 0)
lv = synth.make_pyramid(3, 192, 256, np.float32)
def devout(d):
    t = lambda a: None if a is None else torch.from_numpy(a).to(dev)
    return warp_cost_volume(t(d["c1"]), t(d["c2"]), t(d["flow"]), 4, d["scaler"]).cpu().numpy()
hp = pyramid.HostPyramid()
for sel in ([0], [1], [1, 2], [0, 1, 2, 3, 4]):
    sub = [lv[i] for i in sel]
    outs = [np.full(d["c1"].shape[:3] + (81,), -5.0, np.float32) for d in sub]
    hp.run(sub, outs, 4)
    for d, o in zip(sub, outs):
        ref = devout(d)
        nf = devout({**d, "flow": None})
        print("levels", sel, "lvl", d["lvl"], "diff vs device", float(np.abs(o - ref).max()),
              "vs noflow", float(np.abs(o - nf).max()), "untouched", int((o == -5).sum()))
