"""Summarise an .ncu-rep: key metrics + stall samples aggregated by code region.  usage: ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys, re, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[-1]
for h, v in zip(hdr, vals):
    if h in ('Kernel Name', 'Block Size', 'Grid Size'):
        print(f"{h} = {v}")
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active']
for h, u, v in zip(hdr, units, vals):
    if h in keys:
        print(f"{h} [{u}] = {v}")
for h, u, v in zip(hdr, units, vals):
    if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
        try:
            if float(v) > 0.05: print(f"  stall {h[34:-28]:28s} {float(v):.3f}")
        except ValueError: pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, '# Samples') for r in data)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print('instructions', len(data), 'samples', tot)
B = int(sys.argv[2]) if len(sys.argv) > 2 else 200
for s in range(0, len(data), B):
    chunk = data[s:s + B]
    smp = sum(f(r, '# Samples') for r in chunk)
    if smp < tot * 0.01: continue
    st = {k: sum(f(r, k) for r in chunk) for k in stalls}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
    ex = sum(f(r, 'Instructions Executed') for r in chunk)
    kinds = set()
    for r in chunk:
        o = r[ix['Source']]
        for k in ('FFMA', 'LDG', 'STG', 'SYNCS', 'LDS', 'STS', 'SHFL', 'LDL', 'ATOMS', 'BAR', 'HFMA2'):
            if k in o: kinds.add(k)
    print(f'{s:5d}-{s+B:5d} samples {100*smp/tot:5.1f}% exec {ex/1e6:7.2f}M top {[(k[6:], int(v)) for k, v in top]} {sorted(kinds)}')
