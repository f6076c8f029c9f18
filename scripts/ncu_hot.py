"""Top stalled instructions of a `--page source --csv` dump (gz).  usage: ncu_hot.py file.source.csv.gz [N]"""
import csv, gzip, io, sys
rows = list(csv.reader(io.TextIOWrapper(gzip.open(sys.argv[1]))))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, '# Samples') for r in data)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print('kernel', rows[0][1][:120]); print('instructions', len(data), 'samples', tot)
order = sorted(range(len(data)), key=lambda i: -f(data[i], '# Samples'))[:n]
for i in sorted(order):
    r = data[i]
    st = sorted(((k[6:], int(f(r, k))) for k in stalls), key=lambda kv: -kv[1])[:2]
    print(f"{i:5d} {100 * f(r, '# Samples') / tot:5.1f}% exec {f(r, 'Instructions Executed') / 1e3:8.0f}k  {r[ix['Source']][:70]:70s} {st}")
