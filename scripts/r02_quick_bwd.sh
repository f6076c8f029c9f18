# quick backward check: parity on a few shapes, then timings of every level
set -x
mkdir -p gpurun_out
for args in "1 10 40 32" "2 30 64 32 stress" "1 28 64 64" "1 14 32 96 none" "2 9 33 32 stress" "1 37 70 64 stress" "4 28 64 96 smooth" "3 7 16 128 smooth"; do
  timeout 60 python scripts/check_bwd_one.py $args || { echo "check_bwd_one $args FAILED rc=$?"; exit 1; }
done
PWC_BWD_ONLY_FUSED=1 timeout 200 python scripts/time_bwd_levels.py 32 fp32 2>&1 | tee gpurun_out/quick_time_bwd_fp32.log
