# first GPU pass of round 2: parity of the fused backward, then timings
set -x
mkdir -p gpurun_out; rm -f gpurun_out/parity_errors.txt
for args in "1 10 40 32" "2 30 64 32 stress" "1 28 64 64" "1 14 32 96 none" "2 9 33 32"; do
  timeout 120 python scripts/check_bwd_one.py $args || { echo "check_bwd_one $args FAILED rc=$?"; exit 1; }
done
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "backward or bwd or wrappers or materialised" > gpurun_out/r02_pytest_bwd.log 2>&1; echo "pytest bwd rc=$?"
tail -30 gpurun_out/r02_pytest_bwd.log
timeout 600 python scripts/time_bwd_levels.py 32 fp32 > gpurun_out/r02_time_bwd_fp32.log 2>&1; echo "time rc=$?"; cat gpurun_out/r02_time_bwd_fp32.log
timeout 600 python scripts/time_bwd_levels.py 32 fp16 > gpurun_out/r02_time_bwd_fp16.log 2>&1; echo "time rc=$?"; cat gpurun_out/r02_time_bwd_fp16.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_all.log 2>&1; echo "pytest all rc=$?"
tail -40 gpurun_out/r02_pytest_all.log
