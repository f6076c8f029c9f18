# twenty-ninth GPU pass: pipelined persistent band GEMM (mma_tile 2) against the resident-halo kernel (mma_tile 1)
set -x
mkdir -p gpurun_out; O=gpurun_out
timeout 300 python -m pytest tests -q -m gpu -k "band_gemm" 2>&1 | tail -6
for shp in "32 112 256 32" "32 56 128 64" "4 112 256 32" "2 40 72 32"; do
  for mt in 1 2; do
    timeout 120 python - $shp $mt <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from tfoptflow_b200 import _ffi
_ffi.set_option("mma_tile", int(sys.argv[5]))
print("mma_tile", sys.argv[5], end=" ")
sys.argv = ["check_fwd_impl.py", "4"] + sys.argv[1:5] + ["fp16", "smooth"]
exec(open("scripts/check_fwd_impl.py").read())
PY
  done
done 2>&1 | grep "fwd_impl 4" | sed -E 's/max err.*\); //'
