mkdir -p gpurun_out; O=gpurun_out
for a in "5 0" "-3 0" "20 0" "5 1" "-3 1"; do echo "== load $a"; timeout 60 scripts/ubench/tma_probe $a 2>&1 | grep -v encode; done | tee $O/r02k_tma_probe2.log
