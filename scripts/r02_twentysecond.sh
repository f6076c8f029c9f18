# twenty-second GPU pass: resident-halo band GEMM for search_range 4 (option mma_tile 1) against the 8x32-tile kernel
set -x
for shp in "32 112 256 32" "32 56 128 64" "2 30 52 32" "1 28 40 64"; do
  for mt in 0 1; do
    PWC_OPT_MMA_TILE=$mt timeout 120 python - $shp $mt <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from tfoptflow_b200 import _ffi
_ffi.set_option("mma_tile", int(sys.argv[5]))
sys.argv = ["check_fwd_impl.py", "4"] + sys.argv[1:5] + ["fp16", "stress" if int(sys.argv[1]) < 8 else "smooth"]
exec(open("scripts/check_fwd_impl.py").read())
PY
  done
done 2>&1 | grep "fwd_impl 4"
