# twenty-sixth GPU pass: K1 builders fetch g / out through TMA-staged units (bwd_tma_rows), 3 builders + 9 gather warps
set -x
mkdir -p gpurun_out; O=gpurun_out
for args in "1 10 64 32" "2 30 64 32 stress" "1 28 64 64" "2 9 33 32 stress" "3 7 32 128 smooth" "1 12 96 16 stress" "2 7 32 196 none"; do
  timeout 60 python scripts/check_bwd_one.py $args > $O/chk.log 2>&1 || { echo "check_bwd_one $args FAILED rc=$?"; tail -8 $O/chk.log; exit 1; }
done
echo "parity ok"
for t in 1 0; do
  PWC_BWD_ONLY_FUSED=1 timeout 200 python scripts/time_bwd_levels.py 32 fp32 bwd_tma_rows=$t 2>&1 | sed "s/^/[tma_rows $t] /"
done | tee $O/r02aa_time_bwd.log
PWC_LEVELS=3,2 PWC_BWD_ONLY_FUSED=1 timeout 200 python scripts/time_bwd_levels.py 32 fp16 2>&1 | sed "s/^/[fp16 tma_rows 1] /"
timeout 600 python -m pytest tests -q -m gpu -k "bwd or backward or grad or cfg1 or smoke or options" 2>&1 | tail -4
