# eighteenth GPU pass: K2 scatter warps load the c2 corners half a row at a time (16 LDG.128 in flight) instead of a quarter
set -x
for v in s4; do
  PWC_B200_LIB=$PWD/tfoptflow_b200/_lib/libpwc_$v.so timeout 60 python scripts/check_bwd_one.py 2 30 64 32 stress | tail -3
  PWC_B200_LIB=$PWD/tfoptflow_b200/_lib/libpwc_$v.so PWC_LEVELS=3,2 PWC_BWD_ONLY_FUSED=1 timeout 100 python scripts/time_bwd_levels.py 32 fp32 2>&1 | grep level | sed "s/^/[$v] /"
done
PWC_LEVELS=3,2 PWC_BWD_ONLY_FUSED=1 timeout 100 python scripts/time_bwd_levels.py 32 fp32 2>&1 | grep level | sed "s/^/[default] /"
