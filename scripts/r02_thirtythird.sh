# thirty-third GPU pass: band GEMM on the short 16-bit levels (C = 128, H = 14; C = 96 / 64 with few rows)
set -x
for shp in "32 14 32 128" "8 12 16 128" "2 14 32 128"; do
  timeout 120 python scripts/check_fwd_impl.py 4 $shp fp16 2>&1 | sed -E 's/max err ([0-9.e+-]+) = ([0-9.e+-]+) x rms.*\); /err \2 x rms; /'
done
timeout 120 python scripts/check_fwd_impl.py 4 32 14 32 128 fp16 none 2>&1 | sed -E 's/max err ([0-9.e+-]+) = ([0-9.e+-]+) x rms.*\); /err \2 x rms; /'
