# where does the fused backward's time go: skip one piece at a time (results wrong, timing only), level 2 and 3
set -x
for d in 0 1 2 4 8 16 6 7 14 15 31; do
  PWC_LEVELS=3,2 PWC_BWD_ONLY_FUSED=1 timeout 100 python scripts/time_bwd_levels.py 32 fp32 bwd_debug=$d 2>&1 | grep "level" | sed "s/^/[dbg $d] /"
done
