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
H, W, C = [int(x) for x in os.environ.get("SHAPE", "112,256,32").split(",")]
d = synth.make_level(B, H, W, C, 1971, np.float32, "smooth")
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
names = {1: "P slab start", 2: "P A issued", 3: "P coords done", 4: "P emptyB ok", 5: "P slab full", 8: "C grab", 9: "C waits ok", 10: "C corr done", 11: "C epi done", 12: "S store start", 13: "C refill done", 14: "C pre-fence", 15: "C post-fence"}
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
# per-task phase durations of one warp (WARP env), to spot which tasks are slow
if os.environ.get("WARP"):
    w0 = int(os.environ["WARP"])
    seq = [(t - t0, c, v) for t, w, c, v in ev if w == w0]
    cur = {}
    for t, c, v in seq:
        if c == 8: cur = {"t": v, 8: t}
        elif c in (9, 10, 11) and cur: cur[c] = t
        if c == 11 and all(k in cur for k in (8, 9, 10, 11)):
            print(f"task {cur['t']:4d}: wait {cur[9]-cur[8]:6d} corr {cur[10]-cur[9]:6d} epi {cur[11]-cur[10]:6d}")
