import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tfoptflow_b200.csrc import build as b
V = {
    "v0_ld4_r7": ["-DPWC_STREAM_LD4=ld4"],
    "v1_stream_r7": [],
    "v2_stream_r6": ["-DPWC_STREAM_NSLAB=6", "-DPWC_STREAM_NAP=11"],
    "v3_ld4_r6": ["-DPWC_STREAM_LD4=ld4", "-DPWC_STREAM_NSLAB=6", "-DPWC_STREAM_NAP=11"],
}
os.makedirs(os.path.join(b.LIB_DIR, "variants"), exist_ok=True)
for name, defs in V.items():
    if len(sys.argv) > 1 and name not in sys.argv[1:]: continue
    print(b.build(force=True, out=os.path.join(b.LIB_DIR, "variants", f"lib_{name}.so"), extra_defs=defs), flush=True)
