# twenty-fourth GPU pass: ncu of the resident-halo band GEMM (fp16 level 2), launch list of the fp16 backward
set -x
mkdir -p gpurun_out; O=gpurun_out
bash scripts/ncu_capture.sh r02y_band_L2 mma16_band 3 -- python scripts/check_fwd_impl.py 4 32 112 256 32 fp16 > /dev/null
sed -n 1,60p $O/r02y_band_L2.summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r02y_launches_bwd_fp16.csv \
   env PWC_LEVELS=2 PWC_BWD_ONLY_FUSED=1 python scripts/time_bwd_levels.py 32 fp16 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02y_launches_bwd_fp16.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
for r in rows[-8:]: print(r[ki][:90], r[vi])
PY
