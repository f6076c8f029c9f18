# sixteenth GPU pass: software-pipelined gather in K1 (depth 1 / 2 / 3) and builder : gather warp splits
set -x
mkdir -p gpurun_out; O=gpurun_out
for args in "1 10 40 32" "2 30 64 32 stress" "1 28 64 64" "2 9 33 32 stress" "3 7 16 128 smooth"; do
  timeout 60 python scripts/check_bwd_one.py $args > $O/chk.log 2>&1 || { echo "check_bwd_one $args FAILED rc=$?"; tail -8 $O/chk.log; exit 1; }
done
echo "parity ok"
PWC_BWD_ONLY_FUSED=1 timeout 200 python scripts/time_bwd_levels.py 32 fp32 2>&1 | tee $O/r02r_time_bwd_fp32.log
for v in g1 g3 a5g2 a3g2; do
  PWC_B200_LIB=$PWD/tfoptflow_b200/_lib/libpwc_$v.so PWC_LEVELS=3,2 PWC_BWD_ONLY_FUSED=1 timeout 100 python scripts/time_bwd_levels.py 32 fp32 2>&1 | grep level | sed "s/^/[$v] /"
done 2>&1 | grep "^\[" | tee $O/r02r_variants.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:bwd_fused -s 6 -c 2 --csv --log-file $O/r02r_dram.csv \
     env PWC_LEVELS=2 PWC_BWD_ONLY_FUSED=1 python scripts/time_bwd_levels.py 32 fp32 > /dev/null 2>&1
python - <<'PY'
import csv,sys
rows=[r for r in csv.reader(open('gpurun_out/r02r_dram.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
for r in rows[1:]: print(r[ki][:40], r[mi], r[vi], r[ui])
PY
