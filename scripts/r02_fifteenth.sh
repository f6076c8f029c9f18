# fifteenth GPU pass: bwd_prefetch 1 by default, scatter targets zeroed between K1 and K2, K1 warp-split variants
set -x
mkdir -p gpurun_out; O=gpurun_out
for args in "1 10 40 32" "2 30 64 32 stress" "1 28 64 64" "2 9 33 32 stress" "3 7 16 128 smooth"; do
  timeout 60 python scripts/check_bwd_one.py $args > $O/chk.log 2>&1 || { echo "check_bwd_one $args FAILED rc=$?"; tail -8 $O/chk.log; exit 1; }
done
echo "parity ok"
PWC_BWD_ONLY_FUSED=1 timeout 200 python scripts/time_bwd_levels.py 32 fp32 2>&1 | tee $O/r02p_time_bwd_fp32.log
for v in a5 a7 a8; do
  PWC_B200_LIB=$PWD/tfoptflow_b200/_lib/libpwc_$v.so PWC_LEVELS=3,2 PWC_BWD_ONLY_FUSED=1 timeout 100 python scripts/time_bwd_levels.py 32 fp32 2>&1 | grep level | sed "s/^/[$v] /"
done 2>&1 | grep "^\[" | tee $O/r02p_variants.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:bwd_fused -s 6 -c 2 --csv --log-file $O/r02p_dram.csv \
     env PWC_LEVELS=2 PWC_BWD_ONLY_FUSED=1 python scripts/time_bwd_levels.py 32 fp32 > /dev/null 2>&1
python - <<'PY'
import csv,sys
rows=[r for r in csv.reader(open('gpurun_out/r02p_dram.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
for r in rows[1:]: print(r[ki][:40], r[mi], r[vi], r[ui])
PY
for d in 1 2 4 8 16; do
  PWC_LEVELS=2 PWC_BWD_ONLY_FUSED=1 timeout 100 python scripts/time_bwd_levels.py 32 fp32 bwd_debug=$d 2>&1 | grep "level" | sed "s/^/[dbg $d] /"
done 2>&1 | grep dbg | tee $O/r02p_bwd_dbg.log
