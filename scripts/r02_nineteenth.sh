# nineteenth GPU pass: fused backward with a partial last channel slice (any C % 4 == 0): parity and timing vs the v1 kernels
set -x
mkdir -p gpurun_out; O=gpurun_out
for args in "1 10 40 16" "2 9 33 20 stress" "1 7 16 196 none" "2 7 16 196 smooth" "1 12 20 100 stress" "1 6 8 4 smooth" "1 30 64 36" "1 28 64 64"; do
  timeout 60 python scripts/check_bwd_one.py $args > $O/chk.log 2>&1 || { echo "check_bwd_one $args FAILED rc=$?"; tail -8 $O/chk.log; exit 1; }
done
echo "parity ok"
timeout 100 python scripts/time_bwd_shape.py 8 192 256 16
timeout 100 python scripts/time_bwd_shape.py 32 7 16 196 noflow
timeout 100 python scripts/time_bwd_shape.py 8 6 8 196 noflow
timeout 100 python scripts/time_bwd_shape.py 8 6 8 196
timeout 100 python scripts/time_bwd_shape.py 1 6 8 196 noflow
timeout 100 python scripts/time_bwd_shape.py 32 14 32 128
timeout 600 python -m pytest tests -q -m gpu -k "bwd or backward or grad or cfg1 or smoke or fp16 or bf16" 2>&1 | tail -4
