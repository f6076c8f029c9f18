"""Measures pinned H2D / D2H bandwidth alone and concurrently (what bounds the host-buffer path)."""
import time, torch
n = 400 * 1024 * 1024
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, chunks=1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c = n // chunks
    for i in range(chunks):
        if h2d:
            with torch.cuda.stream(s1): d_in[i*c:(i+1)*c].copy_(h_in[i*c:(i+1)*c], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out[i*c:(i+1)*c].copy_(d_out[i*c:(i+1)*c], non_blocking=True)
    torch.cuda.synchronize(); return time.perf_counter() - t0
for name, a in (("H2D", (True, False)), ("D2H", (False, True)), ("both", (True, True)), ("both x16 chunks", (True, True, 16))):
    run(*a); t = min(run(*a) for _ in range(3))
    print(f"{name:16s}: {t*1e3:7.2f} ms  {n/t/1e9:6.1f} GB/s per direction")
