# twenty-fifth GPU pass: band GEMM variants (gather depth, unroll of the displacement-row loop)
set -x
for v in b200 fb3 dyu3 fb3dyu3; do
  for shp in "32 112 256 32" "32 56 128 64" "32 28 64 96"; do
    PWC_B200_LIB=$PWD/tfoptflow_b200/_lib/libpwc_$v.so timeout 120 python scripts/check_fwd_impl.py 4 $shp fp16 2>&1 | grep "fwd_impl 4" | sed "s/^/[$v] /"
  done
done
