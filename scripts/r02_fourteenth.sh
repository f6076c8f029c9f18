# fourteenth GPU pass: L2 prefetch distance of the backward producers and of the streaming forward vs time and DRAM reads
set -x
mkdir -p gpurun_out; O=gpurun_out
for pf in 0 1 2 3 4; do
  PWC_LEVELS=3,2 PWC_BWD_ONLY_FUSED=1 timeout 100 python scripts/time_bwd_levels.py 32 fp32 bwd_prefetch=$pf 2>&1 | grep "level" | sed "s/^/[bwd_prefetch $pf] /"
done 2>&1 | grep "^\[" | tee $O/r02o_bwd_prefetch.log
for pf in 0 1 2; do
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:bwd_fused -s 6 -c 2 --csv --log-file $O/r02o_dram_pf$pf.csv \
     env PWC_LEVELS=2 PWC_BWD_ONLY_FUSED=1 python scripts/time_bwd_levels.py 32 fp32 bwd_prefetch=$pf > /dev/null 2>&1
  python - $pf <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(f'gpurun_out/r02o_dram_pf{sys.argv[1]}.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
for r in rows[1:]: print('pf', sys.argv[1], r[ki][:40], r[mi], r[vi], r[ui])
PY
done 2>&1 | grep "^pf" | tee $O/r02o_bwd_dram.log
for pf in 0 1 2 4 8; do
  timeout 100 python scripts/time_level.py 32 3 stream_prefetch=$pf 2>&1 | tail -1
done | tee $O/r02o_fwd_prefetch.log
for pf in 0 2 4; do
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:fwd_stream -s 6 -c 1 --csv --log-file $O/r02o_fwd_dram_pf$pf.csv \
     python scripts/time_level.py 32 3 stream_prefetch=$pf > /dev/null 2>&1
  python - $pf <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(f'gpurun_out/r02o_fwd_dram_pf{sys.argv[1]}.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
for r in rows[1:]: print('fwd pf', sys.argv[1], r[ki][:40], r[mi], r[vi], r[ui])
PY
done 2>&1 | grep "^fwd pf" | tee $O/r02o_fwd_dram.log
