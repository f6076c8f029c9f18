# thirtieth GPU pass: band GEMM gather with packed corners (16 loads in flight per thread)
set -x
mkdir -p gpurun_out; O=gpurun_out
timeout 300 python -m pytest tests -q -m gpu -k "band_gemm or fp16 or golden" 2>&1 | tail -4
for shp in "32 112 256 32" "32 56 128 64" "32 28 64 96"; do
  timeout 120 python scripts/check_fwd_impl.py 4 $shp fp16 2>&1 | grep "fwd_impl 4" | sed -E 's/max err.*\); //'
done
R=8 timeout 120 python scripts/check_fwd_impl.py 4 16 272 480 32 fp16 smooth 2>&1 | grep "fwd_impl 4" | sed -E 's/max err.*\); //'
R=8 timeout 120 python scripts/check_fwd_impl.py 4 16 136 240 64 fp16 smooth 2>&1 | grep "fwd_impl 4" | sed -E 's/max err.*\); //'
