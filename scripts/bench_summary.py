"""Reads bench.py output on stdin, echoes the per-level lines and a compact view of the JSON line."""
import json, sys
for l in sys.stdin:
    l = l.strip()
    if l.startswith("{"):
        j = json.loads(l)
        r = j.get("roofline") or {}
        print({k: j.get(k) for k in ("value", "ms_per_step", "gpu_launches")})
        print("roofline", {k: r.get(k) for k in ("achieved", "frac", "kernel_ms", "share_of_step", "traffic")})
        print("e2e", j.get("e2e"))
        print("cpu", j.get("cpu_baseline"))
        print("clocks", j.get("clocks"))
    elif l:
        print(l)
