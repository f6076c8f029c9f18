# thirty-first GPU pass: gather depth variants of the band GEMM
set -x
for v in b200 f5 f4u3; do
  for shp in "32 112 256 32" "32 56 128 64" "32 28 64 96"; do
    PWC_B200_LIB=$PWD/tfoptflow_b200/_lib/libpwc_$v.so timeout 120 python scripts/check_fwd_impl.py 4 $shp fp16 2>&1 | grep "fwd_impl 4" | sed -E 's/max err.*\); //' | sed "s/^/[$v] /"
  done
  PWC_B200_LIB=$PWD/tfoptflow_b200/_lib/libpwc_$v.so R=8 timeout 120 python scripts/check_fwd_impl.py 4 16 272 480 32 fp16 smooth 2>&1 | grep "fwd_impl 4" | sed -E 's/max err.*\); //' | sed "s/^/[$v r8] /"
  PWC_B200_LIB=$PWD/tfoptflow_b200/_lib/libpwc_$v.so R=8 timeout 120 python scripts/check_fwd_impl.py 4 16 136 240 64 fp16 smooth 2>&1 | grep "fwd_impl 4" | sed -E 's/max err.*\); //' | sed "s/^/[$v r8] /"
done
