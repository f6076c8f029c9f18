# twelfth GPU pass: K2 A rows by aligned tensor loads + remainder boxes + in-place residual shift
set -x
mkdir -p gpurun_out; O=gpurun_out
for args in "1 10 40 32" "2 30 64 32 stress" "1 28 64 64" "1 14 32 96 none" "2 9 33 32 stress" "1 37 70 64 stress" "3 7 16 128 smooth" "1 6 20 32 stress" "2 3 35 32 stress"; do
  timeout 60 python scripts/check_bwd_one.py $args > $O/chk.log 2>&1 || { echo "check_bwd_one $args FAILED rc=$?"; tail -8 $O/chk.log; exit 1; }
done
echo "parity ok"
PWC_BWD_ONLY_FUSED=1 timeout 200 python scripts/time_bwd_levels.py 32 fp32 2>&1 | tee $O/r02m_time_bwd_fp32.log
for v in k2a3; do
  PWC_B200_LIB=$PWD/tfoptflow_b200/_lib/libpwc_$v.so PWC_LEVELS=3,2 PWC_BWD_ONLY_FUSED=1 timeout 100 python scripts/time_bwd_levels.py 32 fp32 2>&1 | grep level | sed "s/^/[$v] /"
done 2>&1 | grep "^\[" | tee $O/r02m_variants.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02m_launches_bwd_L2.csv \
   env PWC_LEVELS=2 PWC_BWD_ONLY_FUSED=1 python scripts/time_bwd_levels.py 32 fp32 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02m_launches_bwd_L2.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
for r in rows[-4:]: print(r[ki][:70], r[vi])
PY
for d in 1 2 4 8 16; do
  PWC_LEVELS=2 PWC_BWD_ONLY_FUSED=1 timeout 100 python scripts/time_bwd_levels.py 32 fp32 bwd_debug=$d 2>&1 | grep "level" | sed "s/^/[dbg $d] /"
done 2>&1 | grep dbg | tee $O/r02m_bwd_dbg.log
timeout 600 python -m pytest tests -q -m gpu -k "bwd or backward or grad or cfg1 or smoke" 2>&1 | tail -4
