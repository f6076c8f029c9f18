"""Runs one level-2 launch of the streaming kernel with the event trace on and prints CTA 0's timeline.
usage: trace_stream.py [B] [key=value ...]   (stream_debug bits are OR-ed with 16)"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tfoptflow_b200 import _ffi, synth, pyramid
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dbg = 16
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    if k == "stream_debug": dbg |= int(v)
    else: _ffi.set_option(k, int(v))
_ffi.set_option("fwd_impl", 3)
dev = torch.device("cuda", 0)
d = synth.make_level(B, 112, 256, 32, 1971, np.float32, "smooth")
L = {k: (None if d[k] is None else torch.from_numpy(d[k]).to(dev)) for k in ("c1", "c2", "flow")}
L["scaler"] = 5.0
plan = pyramid.PyramidPlan([L], 4)
for _ in range(3): plan.run()
torch.cuda.synchronize()
_ffi.set_option("stream_debug", dbg)
plan.run(); torch.cuda.synchronize()
NW, TL = 12, 512
buf = (ctypes.c_uint64 * (NW * TL))()
n = _ffi.load().pwc_debug_read_trace(buf, ctypes.sizeof(buf))
a = np.frombuffer(buf, dtype=np.uint64).reshape(NW, TL)
ev = []
for w in range(NW):
    for x in a[w]:
        x = int(x)
        if x == 0: continue
        ev.append((x >> 16, w, (x >> 12) & 0xf, x & 0xfff))
ev.sort()
t0 = ev[0][0]
names = {1: "P slab start", 2: "P A issued", 3: "P coords done", 4: "P emptyB ok", 5: "P slab full", 8: "C grab", 9: "C waits ok", 10: "C corr done", 11: "C epi done"}
lo, hi = int(os.environ.get("T0", 60000)), int(os.environ.get("T1", 75000))
print(f"events={len(ev)} span={ev[-1][0]-t0} cycles")
for t, w, c, v in ev:
    if lo <= t - t0 <= hi:
        print(f"{t - t0:8d} w{w:2d} {names.get(c, c):14s} {v}")
# per-phase statistics
import collections
dur = collections.defaultdict(list)
last = {}
for t, w, c, v in ev:
    if w in last:
        pc, pt = last[w]
        dur[(pc, c)].append(t - pt)
    last[w] = (c, t)
for k in sorted(dur):
    x = np.array(dur[k])
    print(f"{names.get(k[0], k[0]):14s} -> {names.get(k[1], k[1]):14s}: n={len(x):4d} mean={x.mean():8.0f} med={np.median(x):8.0f} max={x.max():8d}")
